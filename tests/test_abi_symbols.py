"""CPU: the C-ABI library builds, loads and exports every symbol include/l3dpp_b200.h declares;
without a CUDA device the product refuses to run (no CPU fallback)."""
import ctypes as C
import os

import pytest

from conftest import has_gpu


def test_library_exports_every_declared_symbol(api):
    api.build()
    assert os.path.exists(api.LIB_PATH)
    names = api.declared_symbols()
    assert len(names) >= 25, names
    L = C.CDLL(api.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_error_strings(api):
    L = api.lib()
    assert b"sm_100a" in L.l3d_version()
    assert isinstance(L.l3d_last_error(), bytes)


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(api):
    with pytest.raises(api.L3DError) as e:
        api.Context()
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_host_clustering_matches_oracle_kat(api, oracle):
    import numpy as np
    rng = np.random.default_rng(5)
    n = 60
    ij = rng.integers(0, n, size=(400, 2)).astype(np.int32)
    ij = ij[ij[:, 0] != ij[:, 1]]
    w = (0.5 + 0.5 * rng.random(len(ij))).astype(np.float32)
    w[::7] = w[3]  # ties exercise the stable sort
    both = np.concatenate([ij, ij[:, ::-1]], axis=1).reshape(-1, 2)  # (i,j),(j,i) like A_
    wb = np.repeat(w, 2)
    got = api.cluster_edges(both, wb, n)
    exp = np.zeros(n, dtype=np.int32)
    m = oracle.lib().orc_kat_cluster(both.ctypes.data, wb.ctypes.data, len(wb), n, exp.ctypes.data)
    assert m == n
    assert (got == exp).all()


def test_header_is_plain_c():
    """include/l3dpp_b200.h is a C header (the cgo / JNI / ctypes side of the boundary): it parses as C99 and as C++11."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "l3dpp_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])


def test_ctypes_mirrors_have_the_c_layout(api, tmp_path):
    """The ctypes structures of api.py against sizeof / offsetof of the C header, compiled with gcc: a field added on
    one side only (l3d_counts.pair_tests_run was the last one) shifts everything behind it."""
    import ctypes
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "l3dpp_b200.h"
int main(void) {
    printf("%zu %zu %zu\\n", sizeof(l3d_view), sizeof(l3d_params), sizeof(l3d_counts));
    printf("%zu %zu %zu %zu\\n", offsetof(l3d_counts, pair_tests_run), offsetof(l3d_counts, num_views),
           offsetof(l3d_counts, gpu_launches), offsetof(l3d_params, shard_world));
    return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes, offs = [int(x) for x in out[:3]], [int(x) for x in out[3:]]
    assert sizes == [ctypes.sizeof(api.View), ctypes.sizeof(api.Params), ctypes.sizeof(api.Counts)]
    assert offs == [api.Counts.pair_tests_run.offset, api.Counts.num_views.offset, api.Counts.gpu_launches.offset,
                    api.Params.shard_world.offset]
