"""CPU: the reference-side bindings compile against the C ABI header (stand-alone stand-ins of the
reference types) and link against libl3dpp_b200.so."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3dline-slam_b200")


def test_cudawrapper_shim_compiles(tmp_path, api):
    api.build()
    obj = tmp_path / "shim.o"
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                           os.path.join(PKG, "shim", "cudawrapper_b200.cpp"), "-o", str(obj)])
    syms = subprocess.check_output(["nm", "-C", str(obj)], text=True)
    for s in ("L3DPP::match_lines_GPU(", "L3DPP::score_matches_GPU(", "L3DPP::match_lines_GPU_f64(",
              "L3DPP::find_collinear_segments_GPU(", "L3DPP::replicator_dynamics_diffusion_GPU("):
        assert s in syms, s


def test_cpp_line3d_mirror_links(tmp_path, api):
    api.build()
    src = tmp_path / "t.cpp"
    src.write_text('#include "line3d_b200.hpp"\n'
                   'int main() { L3DPP_B200::Line3D l("", false, 640); l.matchImages();\n'
                   '  L3DPP_B200::Line3DStream s("", false, 640); s.beginCycle(); s.deleteImage(3); s.matchImages();\n'
                   '  s.reconstruct3Dlines(); return (int)l.numImages(); }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-Wall", "-I", os.path.join(ROOT, "include"), "-I",
                           os.path.join(PKG, "host"), str(src), "-L", PKG, "-ll3dpp_b200", "-Wl,-rpath," + PKG, "-o",
                           str(exe)])
    # without a GPU the context cannot be created: the mirror prints the error and keeps going
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0
