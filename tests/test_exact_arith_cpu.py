"""The exact replacements the kernels use for IEEE divisions / conversions (csrc/ot_rowmath.cuh), restated in C and checked on the
CPU against the plain expressions of the reference arithmetic: tools/check_quant_bits.c (quant4_pack: rint and the
int8 byte from the bit pattern of q + 1.5*2^23) and a strided run of tools/check_div127.c's sequence (x / 127 as three FMAs).
The exhaustive runs (every float of the range) are documented in the tools' headers; here a bounded sample keeps the CPU suite fast."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GCC = shutil.which("gcc")

pytestmark = pytest.mark.skipif(GCC is None, reason="gcc not available")


def _build(tmp_path, src, extra_src=None, flags=()):
    exe = str(tmp_path / "check")
    path = os.path.join(ROOT, "tools", src) if extra_src is None else str(tmp_path / src)
    if extra_src is not None:
        with open(path, "w") as f:
            f.write(extra_src)
    subprocess.run([GCC, "-O2", "-ffp-contract=off", *flags, "-o", exe, path, "-lm"], check=True)
    return exe


def test_quant_bits_matches_rint_of_true_division(tmp_path):
    exe = _build(tmp_path, "check_quant_bits.c")
    res = subprocess.run([exe, "3000000"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 mismatches" in res.stdout


@pytest.mark.parametrize("flags", [(), ("-DSEQ4",)])
def test_branch_free_exact_quotient_of_the_requant_epilogues(tmp_path, flags):
    """tools/check_div_exact.c: the FMA-residual quotient of the fused RowQuant epilogues (ot_gemm_wres.cu, ot_attention_tc.cu) equals the
    IEEE division bit for bit on random and adversarial (near half-integer) cases -- the two-step form the kernels use and the
    four-operation form (-DSEQ4) that was measured against it; a bounded sample of the 8e9-case runs quoted in the tool's header."""
    exe = _build(tmp_path, "check_div_exact.c", flags=flags)
    res = subprocess.run([exe, "30000000"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 mismatches" in res.stdout


def test_div127_and_div511_sequences_strided(tmp_path):
    src = r'''
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static float seq(float x, float D) {
  const float r = 1.0f / D;
  const float r2 = fmaf(fmaf(-D, r, 1.0f), r, r);
  const float q0 = x * r2;
  return fmaf(fmaf(-D, q0, x), r2, q0);
}
static uint64_t run(float D, float lo, float hi, uint32_t stride) {
  uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4);
  uint64_t bad = 0;
  for (uint64_t u = a; u <= b; u += stride) {
    uint32_t v = (uint32_t)u; float x; memcpy(&x, &v, 4);
    float t = x / D, s = seq(x, D);
    if (memcmp(&t, &s, 4) != 0) ++bad;
  }
  return bad;
}
int main(void) {
  uint64_t bad = run(127.0f, 1e-5f, 1e30f, 197) + run(511.0f, 1e-37f, 1e37f, 389);
  printf("mismatches %llu\n", (unsigned long long)bad);
  return bad != 0;
}
'''
    exe = _build(tmp_path, "div_strided.c", src)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
