// TEST INFRASTRUCTURE: the host function the device path uses for the penetration scan of FSI::apply_contact_model
// (openifem_b200/csrc/contact.h) behind a C entry point, so that tests/test_contact_scan_cpu.py can run the very same code
// against the oracle without a GPU. Built with g++ together with the (host-only) mesh.cpp / fe_tables.cpp of the product.
#include "../../openifem_b200/csrc/contact.h"

typedef double (*point_fn)(const double *);

extern "C" int cpu_contact_scan(int dim, int degree, int n_bfaces, const int *boundary_faces, const int *cell_nodes, int npc,
                                const double *coords, const double *u, long long n_dofs, point_fn criterion, const double *direction,
                                double multiplier, double *rows)
{
  const ifem::ContactScan scan(dim, degree);
  return scan.run(n_bfaces, boundary_faces, cell_nodes, npc, coords, u, n_dofs, [criterion](const double *p) { return criterion(p); },
                  direction, multiplier, rows)
           ? 1
           : 0;
}
