"""On-disk formats either side of the hot path (openifem_b200/csrc/output.h), host side, no device:
.vtu files parse as VTK XML and carry the reference's field names with the right values at the right vertices
(FluidSolver::output_results, source/mpi_fluid_solver.cpp:491-579), the .pvd collection has Utils::PVDWriter's layout
(source/utilities.cpp:38-81), and the solid checkpoint streams follow deal.II's Vector<double>::block_write
("<size>\\n[" + raw doubles + "]", used by source/mpi_shared_solid_solver.cpp:452-571)."""
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

import openifem_b200 as ifem
from oracle import fem


def _arrays(path):
    root = ET.parse(path).getroot()
    assert root.tag == "VTKFile" and root.attrib["type"] == "UnstructuredGrid"
    piece = root.find("UnstructuredGrid/Piece")
    out = {"n_points": int(piece.attrib["NumberOfPoints"]), "n_cells": int(piece.attrib["NumberOfCells"])}
    for sec in ("Points", "Cells", "PointData", "CellData"):
        for a in piece.find(sec).findall("DataArray"):
            v = np.array(a.text.split(), dtype=np.float64)
            nc = int(a.attrib.get("NumberOfComponents", 1))
            out[sec + "/" + a.attrib.get("Name", "points")] = v.reshape(-1, nc) if nc > 1 else v
    return out


def test_write_vtu_quads_and_hexes(tmp_path):
    for dim in (2, 3):
        m = fem.BoxMesh((3, 2) if dim == 2 else (2, 2, 3), (0,) * dim, (1.0,) * dim)
        rng = np.random.default_rng(dim)
        vec, sca, cel = rng.uniform(-1, 1, (m.vertices.shape[0], dim)), rng.uniform(-1, 1, m.vertices.shape[0]), rng.uniform(-1, 1, m.n_cells)
        path = str(tmp_path / f"mesh{dim}.vtu")
        ifem.io.write_vtu(path, dim, m.vertices, m.cells, [("velocity", vec), ("pressure", sca)], [("Indicator", cel)])
        a = _arrays(path)
        assert a["n_points"] == m.vertices.shape[0] and a["n_cells"] == m.n_cells
        assert np.array_equal(a["Points/points"][:, :dim], m.vertices) and (dim == 3 or np.all(a["Points/points"][:, 2] == 0))
        assert np.array_equal(a["PointData/velocity"][:, :dim], vec) and a["PointData/velocity"].shape[1] == 3  # padded like deal.II
        assert np.array_equal(a["PointData/pressure"], sca) and np.array_equal(a["CellData/Indicator"], cel)
        assert np.all(a["Cells/types"] == (9 if dim == 2 else 12))
        conn = a["Cells/connectivity"].astype(int).reshape(m.n_cells, -1)
        # VTK order = lexicographic order with vertices 2 <-> 3 (and 6 <-> 7) swapped; the cells must be positively oriented
        perm = [0, 1, 3, 2] if dim == 2 else [0, 1, 3, 2, 4, 5, 7, 6]
        assert np.array_equal(conn, m.cells[:, perm])
        X = m.vertices[conn]
        if dim == 2:
            e1, e2 = X[:, 1] - X[:, 0], X[:, 3] - X[:, 0]
            assert np.all(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0] > 0)
        else:
            e1, e2, e3 = X[:, 1] - X[:, 0], X[:, 3] - X[:, 0], X[:, 4] - X[:, 0]
            assert np.all(np.einsum("ci,ci->c", np.cross(e1, e2), e3) > 0)


def test_fluid_results_file_has_the_reference_fields_at_the_vertices(tmp_path):
    """Q2/Q1 solution sampled at the cell vertices (build_patches(pressure degree = 1)); names of :497-560"""
    for dim in (2, 3):
        reps = (4, 3) if dim == 2 else (2, 3, 2)
        tria = ifem.Triangulation(dim)
        ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, (2.0,) + (1.0,) * (dim - 1), True)
        d = fem.FluidDofs(fem.BoxMesh(reps, (0,) * dim, (2.0,) + (1.0,) * (dim - 1)), 2, 1)
        pts = d.support_points()
        # fields that are functions of the position, so that every written value can be checked at its own point
        present = np.empty(d.n_dofs)
        for c in range(dim):
            present[c: d.n_u: dim] = (c + 1) * pts[c: d.n_u: dim, 0] - 0.5 * pts[c: d.n_u: dim, 1]
        present[d.n_u:] = 3.0 + pts[d.n_u:, 0] * pts[d.n_u:, 1]
        acc = -2.0 * present
        ind = (np.arange(d.mesh.n_cells) % 3 == 0).astype(np.int32)
        stress = np.stack([(k + 1) * d.ucoords[:, 0] + d.ucoords[:, -1] for k in range(dim * dim)])
        ifem.io.fluid_write_results_host(tria, 2, 1, present, acc, ind, stress, str(tmp_path), 7)
        a = _arrays(str(tmp_path / "fluid_000007.proc0000.vtu"))
        assert a["n_points"] == d.mesh.vertices.shape[0] and a["n_cells"] == d.mesh.n_cells
        x = a["Points/points"]
        for c in range(dim):
            assert np.allclose(a["PointData/velocity"][:, c], (c + 1) * x[:, 0] - 0.5 * x[:, 1], atol=1e-14)
            assert np.allclose(a["PointData/fsi_force"][:, c], -2.0 * ((c + 1) * x[:, 0] - 0.5 * x[:, 1]), atol=1e-14)
        assert np.allclose(a["PointData/pressure"], 3.0 + x[:, 0] * x[:, 1], atol=1e-14)
        assert np.allclose(a["PointData/dummy_fsi_force"], -2.0 * (3.0 + x[:, 0] * x[:, 1]), atol=1e-14)
        names = ["Txx", "Txy", "Tyy"] + (["Txz", "Tyz", "Tzz"] if dim == 3 else [])
        comp = {"x": 0, "y": 1, "z": 2}
        for n in names:
            k = comp[n[1]] * dim + comp[n[2]]
            assert np.allclose(a["PointData/" + n], (k + 1) * x[:, 0] + x[:, dim - 1], atol=1e-14), n
        assert np.array_equal(a["CellData/Indicator"], ind) and np.all(a["CellData/subdomain"] == 0)
        # the master record lists the piece and declares the same arrays
        root = ET.parse(str(tmp_path / "fluid_000007.pvtu")).getroot()
        assert root.attrib["type"] == "PUnstructuredGrid"
        assert [p.attrib["Source"] for p in root.iter("Piece")] == ["fluid_000007.proc0000.vtu"]
        declared = [p.attrib["Name"] for p in root.find("PUnstructuredGrid/PPointData")]
        assert declared == ["velocity", "pressure", "fsi_force", "dummy_fsi_force"] + names


def test_pvd_collection_layout(tmp_path):
    path = str(tmp_path / "fluid.pvd")
    ifem.io.write_pvd(path, "fluid_", [0.0, 0.01, 0.02], [0, 1, 2])
    text = open(path).read()
    assert text.startswith('<?xml version="1.0"?>\n<!--\n#This file was generated by OpenIFEM on ')
    assert '<VTKFile type="Collection" version="0.1" ByteOrder="LittleEndian">\n  <Collection>\n' in text
    assert '    <DataSet timestep="0.01" group="" part="0" file="fluid_000001.pvtu"/>\n' in text
    assert text.endswith("  </Collection>\n</VTKFile>\n")
    sets = ET.parse(path).getroot().findall("Collection/DataSet")
    assert [s.attrib["file"] for s in sets] == ["fluid_%06d.pvtu" % k for k in range(3)]
    assert [float(s.attrib["timestep"]) for s in sets] == [0.0, 0.01, 0.02]


def test_block_write_is_dealii_vector_stream(tmp_path):
    v = np.random.default_rng(0).uniform(-1, 1, 37)
    path = str(tmp_path / "000005.solid_checkpoint_displacement")
    ifem.io.block_write(path, v)
    raw = open(path, "rb").read()
    assert raw[:4] == b"37\n[" and raw[-1:] == b"]" and len(raw) == 4 + 37 * 8 + 1
    assert np.array_equal(np.array(struct.unpack("<37d", raw[4:-1])), v)
    assert np.array_equal(ifem.io.block_read(path, 64), v)
