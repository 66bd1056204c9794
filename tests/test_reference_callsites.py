"""The drop-in boundary against the reference's REAL call sites (marker `ref`: needs /root/reference, compile only).

src/density_clustering.cpp is compiled UNMODIFIED with -DUSE_CUDA -- Density::main included, i.e. the six call sites
:113-118, :616-621, :659-663, :716-720, :746-750, :808-814 -- in a scratch copy of the reference's src/ in which only
density_clustering_cuda.hpp is replaced by this repository's forwarding header (what cmake/dcb200.cmake does), with the
reference's own Pops / Neighborhood types.  The object must then link against libdcb200.so with no unresolved
Clustering::Density::CUDA symbol left.  Boost is absent from the image: tests/callsite_shim/ stands in for
variables_map (declaration-level only)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
pytestmark = pytest.mark.ref


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")
def test_reference_call_sites_compile_and_link_against_the_shim(tmp_path):
    src = tmp_path / "src"
    src.mkdir()
    for f in os.listdir(REF):
        if f.endswith((".cpp", ".hpp", ".hxx")) and f != "density_clustering_cuda.hpp":
            os.symlink(os.path.join(REF, f), src / f)
    shutil.copy(os.path.join(ROOT, "include", "dcb200", "reference_tree", "density_clustering_cuda.hpp"), src / "density_clustering_cuda.hpp")
    shutil.copy(os.path.join(ROOT, "oracle", "shim", "config.hpp"), src / "config.hpp")
    flags = ["-std=c++11", "-O1", "-fopenmp", "-fPIC", "-DUSE_CUDA", "-include", "limits", "-include", "cmath", "-include", "set",
             "-I", os.path.join(ROOT, "tests", "callsite_shim"), "-I", os.path.join(ROOT, "include")]
    objs = []
    for tu in ("density_clustering.cpp", "tools.cpp", "logger.cpp"):
        obj = str(tmp_path / (tu + ".o"))
        out = subprocess.run(["g++"] + flags + ["-c", str(src / tu), "-o", obj], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr[-4000:]
        objs.append(obj)
    # the translation unit really went through the CUDA call sites: it references the C ABI, not the reference's .cu objects
    nm = subprocess.run(["nm", "-C", objs[0]], capture_output=True, text=True).stdout
    for sym in ("dcb200_populations", "dcb200_nearest_neighbors", "dcb200_screening_next", "dcb200_screening"):
        assert f" U {sym}" in nm, sym
    assert "Clustering::Density::CUDA::calculate_populations" in nm          # the inline shim functions were instantiated here
    # link: everything except `main`-level pieces of the other sub-modules must resolve against libdcb200.so
    lib = os.path.join(ROOT, "clustering_b200")
    out = subprocess.run(["g++", "-shared", "-fopenmp", "-o", str(tmp_path / "libcallsites.so")] + objs +
                         ["-L", lib, "-ldcb200", "-Wl,--no-undefined", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]
