"""CPU-side checks of product pieces that do not need a GPU: the median-of-9 selection network and
scipy-'reflect' index map shared with the CUDA prep kernel (csrc/median9.h, built here with gcc),
the C-ABI export list, and the descriptor planner."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
from scipy.ndimage import median_filter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "whisper-timestamped_b200", "csrc")


@pytest.fixture(scope="module")
def medlib(tmp_path_factory):
    d = tmp_path_factory.mktemp("med")
    src = d / "med.c"
    src.write_text(f'''
#include "{CSRC}/median9.h"
void median_rows(const float* x, int rows, int n, float* out) {{
    for (int r = 0; r < rows; ++r) for (int c = 0; c < n; ++c) {{
        float p[9];
        for (int k = 0; k < 9; ++k) p[k] = x[r * n + wts_reflect_index(c - 4 + k, n)];
        out[r * n + c] = wts_median9(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
    }}
}}
void median_rows_pair(const float* x, int rows, int n, float* out) {{
    /* same data flow as the CUDA prep kernel: halo-padded row, two outputs per step */
    float xb[4096];
    for (int r = 0; r < rows; ++r) {{
        for (int c = 0; c < n; ++c) xb[c + 4] = x[r * n + c];
        for (int k = 0; k < 4; ++k) {{ xb[3 - k] = xb[4 + wts_reflect_index(-1 - k, n)]; xb[n + 4 + k] = xb[4 + wts_reflect_index(n + k, n)]; }}
        xb[n + 8] = 0.f;
        for (int c = 0; c < n; c += 2) {{
            float m0, m1;
            wts_median9_pair(xb + c, &m0, &m1);
            out[r * n + c] = m0;
            if (c + 1 < n) out[r * n + c + 1] = m1;
        }}
    }}
}}
int reflect_index(int p, int n) {{ return wts_reflect_index(p, n); }}
''')
    so = d / "med.so"
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(so), str(src), "-lm"])
    return ctypes.CDLL(str(so))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 9, 10, 17, 150, 1500])
def test_median9_matches_scipy(medlib, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((7, n)).astype(np.float32)
    x[0, : min(n, 5)] = 1.0                      # ties
    out = np.empty_like(x)
    medlib.median_rows(x.ctypes.data_as(ctypes.c_void_p), 7, n, out.ctypes.data_as(ctypes.c_void_p))
    ref = median_filter(x, (1, 9))               # same call as transcribe.py:1546 (mode='reflect')
    assert np.array_equal(out, ref)
    out2 = np.empty_like(x)
    medlib.median_rows_pair(x.ctypes.data_as(ctypes.c_void_p), 7, n, out2.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(out2, ref)


def test_reflect_index_is_numpy_symmetric(medlib):
    for n in (1, 2, 3, 5, 9, 40):
        padded = np.pad(np.arange(n), (4, 4), mode="symmetric")
        got = [medlib.reflect_index(p, n) for p in range(-4, n + 4)]
        assert got == padded.tolist()


def test_abi_exports_every_declared_symbol():
    """libwts.so loads on a CPU-only box and exports what include/wts.h declares."""
    from whisper_timestamped import _native as nat
    header = open(os.path.join(ROOT, "include", "wts.h")).read()
    declared = set(re.findall(r"\b(wts_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(nat.lib, name), f"{name} declared in include/wts.h but not exported"
    assert set(nat.EXPORTED_SYMBOLS) <= declared
    assert nat.lib.wts_version() >= 100


def test_disfluency_peaks_match_scipy(tmp_path):
    """csrc/peaks.h (what the CUDA disfluency kernel runs) against scipy.signal.find_peaks(width=3, prominence=0.02)
    itself: same 'more than one peak' decision and the same round(left_ips[-1]) on random, plateau and tiny inputs."""
    from scipy.signal import find_peaks
    src = tmp_path / "pk.c"
    src.write_text(f'''
#include "{CSRC}/peaks.h"
int disfluency_left(const float* row, int begin, int n) {{ return wts_disfluency_left(row, begin, n, 0.02, 3.0); }}
''')
    so = tmp_path / "pk.so"
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(so), str(src), "-lm"])
    lib = ctypes.CDLL(str(so))
    lib.disfluency_left.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    rng = np.random.default_rng(5)
    n_multi = 0
    for trial in range(3000):
        n = int(rng.integers(0, 60))
        begin = int(rng.integers(0, 5))
        kind = trial % 4
        if kind == 0:                       # smooth bumps (what attention rows look like)
            t = np.arange(n + begin)
            v = sum(rng.uniform(0.01, 0.3) * np.exp(-((t - rng.uniform(0, n + begin)) / rng.uniform(1.5, 6)) ** 2)
                    for _ in range(int(rng.integers(1, 4))))
            v = np.asarray(v, dtype=np.float64) * np.ones(n + begin)
        elif kind == 1:                     # noise
            v = rng.uniform(0, 0.2, n + begin)
        elif kind == 2:                     # quantised: plateaus and exact ties
            v = np.round(rng.uniform(0, 0.2, n + begin) * 20) / 20
        else:                               # smooth + noise
            t = np.arange(n + begin)
            v = 0.1 * np.sin(t / rng.uniform(1.0, 5.0)) ** 2 + rng.uniform(0, 0.01, n + begin)
        cost = (-np.asarray(v)).astype(np.float32)          # the kernel reads float32 costs and negates them
        x = -cost[begin:begin + n].astype(np.float64)
        peaks, props = find_peaks(x, width=3, prominence=0.02)
        want = int(round(props["left_ips"][-1])) if len(peaks) > 1 else -1
        n_multi += len(peaks) > 1
        got = lib.disfluency_left(cost.ctypes.data, begin, n)
        assert got == want, (trial, n, begin, got, want, x.tolist())
    assert n_multi > 100          # the interesting branch was exercised


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors (whisper_timestamped/_native.py) must have the size and field offsets gcc gives the structs
    of include/wts.h — this is the C-ABI boundary the reference-side binding relies on."""
    import ctypes
    import subprocess
    from whisper_timestamped import _native as nat
    mirrors = {"WtsSegDesc": nat.SegDesc, "WtsGemm": nat.Gemm, "WtsDecodeCfg": nat.DecodeCfg,
               "WtsDecLayer": nat.DecLayer, "WtsDecodeSteps": nat.DecodeSteps}
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "wts.h"', "int main(void) {"]
    for cname, cls in mirrors.items():
        src.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            src.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    src += ["  return 0;", "}"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    seen = 0
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        cname, fname, val = line.split()
        cls = mirrors[cname]
        if fname == "size":
            assert ctypes.sizeof(cls) == int(val), (cname, ctypes.sizeof(cls), val)
        else:
            assert getattr(cls, fname).offset == int(val), (cname, fname, getattr(cls, fname).offset, val)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in mirrors.values())


def test_plan_segments_layout():
    from whisper_timestamped import _native as nat
    from whisper_timestamped.alignment import plan_segments
    plan = plan_segments([(0, 0, None, 24, 10, 300, 0), (1, 5, 40, 33, 0, 90, 50), (0, 30, None, 2, 0, 9, 0)])
    s = plan.segs
    assert s["last_row"].tolist() == [23, 40, 31]
    assert s["cost_off"].tolist() == [0, 7200, 7200 + 33 * 92]       # rows padded to 16 bytes: pitch 92 for F = 90
    assert (s["flags"] & 2).all() and plan.cost_elems == 7200 + 33 * 92 + 2 * 12
    assert s["jumps_off"].tolist() == [0, 25, 59]
    assert plan.jumps_elems == 62 and plan.max_T == 33 and plan.max_F == 300
    assert s["dir_off"][1] == nat.lib.wts_dtw_dir_words(24, 300)
    assert nat.lib.wts_dtw_bnd_doubles(24, 300) == 0 and nat.lib.wts_dtw_bnd_doubles(33, 90) == 92
    assert plan.dtw_order.tolist() == [0, 1, 2]
    with pytest.raises(ValueError):
        plan_segments([(0, 0, None, 0, 0, 5, 0)])


def test_no_cpu_fallback():
    import torch
    from whisper_timestamped import _native as nat
    from whisper_timestamped.alignment import plan_segments, attn_prep
    plan = plan_segments([(0, 0, None, 3, 0, 9, 0)])
    with pytest.raises(nat.WtsError):
        attn_prep(torch.zeros(1, 2, 4, 16), plan)
