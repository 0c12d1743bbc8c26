"""tools/sass_patch.py (the SASS pass of DESIGN.md 4.0) on a small sm_100a cubin: every register-form IMAD.MOV.U32 becomes a MOV,
IMAD.IADD / IMAD.X become IADD3 / IADD3.X, nothing else changes, and the patched cubin still disassembles.  CPU-only (nvcc
cross-compiles; nothing is executed)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

SRC = r"""
#include <cstdint>
__device__ __noinline__ uint64_t f(uint32_t a, uint32_t b, uint64_t c) {
  uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32), t;
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, 0, 0;" : "+r"(lo), "+r"(hi), "=r"(t) : "r"(a), "r"(b));
  return ((uint64_t)(hi + t) << 32) | lo;
}
extern "C" __global__ void k(uint64_t* out, const uint32_t* in, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t acc = in[i];
  for (int j = 0; j < n; j++) acc = f(in[(i + j) & 1023], in[(i + 2 * j) & 1023], acc) + f((uint32_t)acc, in[j & 1023], acc >> 7);
  out[i] = acc;
}
"""


def _sass(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    return [m.group(1).strip() for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", out, re.M)]


@pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None, reason="needs the CUDA toolkit")
def test_sass_pass_rewrites_only_the_moves(tmp_path):
    import sass_patch
    cu, cubin, patched = tmp_path / "k.cu", tmp_path / "k.cubin", tmp_path / "k_patched.cubin"
    cu.write_text(SRC)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-cubin", "-o", str(cubin), str(cu)])
    before = _sass(str(cubin))
    stats = sass_patch.main([str(cubin), str(patched), "--mindist", "5"])
    after = _sass(str(patched))
    assert len(before) == len(after)
    n_mov = 0
    for b, a in zip(before, after):
        if b == a:
            continue
        m = re.match(r"(@!?P\d\s+)?IMAD\.MOV\.U32 (R\d+|RZ), RZ, RZ, (R\d+|RZ)$", b)
        if m:
            assert a == f"{m.group(1) or ''}MOV {m.group(2)}, {m.group(3)}", (b, a)
            n_mov += 1
            continue
        # IMAD.IADD / IMAD.X / the signed IMAD.MOV (= 0 * 0 + Rs) become additions on the ALU pipe
        assert re.match(r"(@!?P\d\s+)?IMAD\.(IADD|X|MOV) ", b) and re.match(r"(@!?P\d\s+)?IADD3(\.X)? ", a), (b, a)
    assert n_mov == stats["mov"] and n_mov > 0
    assert not any(re.search(r"IMAD\.MOV\.U32 (R\d+|RZ), RZ, RZ, (R\d+|RZ)$", a) for a in after)
