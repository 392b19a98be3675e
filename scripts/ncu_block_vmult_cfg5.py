"""One block mat-vec of the SCnsIM system on the config-5 mesh at a reduced scale (for an ncu capture of its four kernels).
    ncu --set full --clock-control none -k regex:bcsr_spmv -c 8 python scripts/ncu_block_vmult_cfg5.py [scale]"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench

import openifem_b200 as ifem

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ifem.init(0)
ftria, _ = bench.fsi_meshes(5, scale, 1)
fluid = ifem.Fluid.MPI.SCnsIM(ftria, ifem.Parameters.AllParameters(bench.fsi_prm_path(5)))
fluid.setup()
fluid.assemble(True)
ms, b = fluid.bench_vmult(2)
print(f"block mat-vec {ms:.4f} ms, {b / 1e9:.3f} GB algorithmic, sizes {fluid.sizes()}")
