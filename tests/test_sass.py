"""Static checks of the built library (no GPU needed: cuobjdump reads the sm_100a cubins in libnfk.so):
the hot kernels use the Blackwell instructions the design claims (tcgen05.mma = UTCHMMA, TMA bulk copies =
UBLKCP, packed fp32x2 = FFMA2/FADD2, TMEM loads = LDTM), the spline element costs the MUFU operations the
roofline analysis counts, and none of them spills registers to local memory."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "normalizingflow_b200", "libnfk.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)),
                                reason="needs the built libnfk.so and cuobjdump")


def _usage():
    out = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    res, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
        elif name and "REG:" in line:
            res[name] = {k: int(v) for k, v in re.findall(r"(REG|STACK|LOCAL):(\d+)", line)}
            name = None
    return res


def _sass(fn):
    return subprocess.run([CUOBJDUMP, "-sass", "-fun", fn, LIB], capture_output=True, text=True, check=True).stdout


def _count(sass, mnemonic):
    return len(re.findall(r"\s" + re.escape(mnemonic) + r"[\s.;]", sass))


FUSED = "_ZN3nfk22nsf_pairs_fused_kernelILi2ELb0ELb0EEEvNS_9FusedArgsE"      # FAST, forward, production (no debug hook)
PAIRS = "_ZN3nfk18rqs_coupling_pairsILi2ELi8ELb0EEEvNS_12CouplingArgsE"


def test_only_sm100a_code_and_no_spills_in_the_hot_kernels():
    archs = subprocess.run([CUOBJDUMP, "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in archs and not re.search(r"sm_(?!100a)\d+", archs), archs
    res = _usage()
    # K = 8 instantiations of the spline kernels (the generic-K ones hold up to 2 x 33 knots and spill by design)
    hot = [k for k in res if any(t in k for t in ("nsf_pairs_fused_kernel", "gemm_ws_kernel", "wgrad_ws_kernel",
                                                  "linear_tf32x3_kernel"))
           or re.search(r"rqs_coupling_(pairs|tiled)ILi\dELi8E", k)]
    assert len(hot) >= 20
    # (the lazy exact-bin path of the FAST spline epilogue costs one gemm_ws instantiation two spilled
    # registers on its cold side: tolerated up to 16 bytes, nothing in the fused layer kernel)
    hot = [k for k in hot if not re.search(r"nsf_pairs_fused_kernelILi\dELb[01]ELb1E", k)]      # debug-hook instantiations: tests only
    hot += [k for k in res if re.search(r"nsf_fused2_kernelILi[12]ELb[01]ELb[01]ELb0E", k)]       # production FAST / HYBRID, both operand modes
    spilled = {k: v for k, v in res.items() if k in hot and (v["LOCAL"] or v["STACK"] > (16 if "gemm_ws_kernelILi2E" in k else 0))}
    assert not spilled, spilled
    # the fused layer kernel shares sub-partition 0 with its control warp: 5 warps -> at most 96 registers
    assert res[FUSED]["REG"] <= 96, res[FUSED]


def test_fused_layer_kernel_instruction_mix():
    s = _sass(FUSED)
    assert _count(s, "UTCHMMA") == 20                 # 4 (GEMM1) + 8 (GEMM2) + 8 (one GEMM3 chunk) tcgen05.mma
    assert _count(s, "UBLKCP") >= 4                   # TMA bulk copies: weights, W3 ring, x rows in, z rows out
    assert _count(s, "LDTM") >= 4                     # tcgen05.ld: two hidden-layer epilogues + the spline chunk
    assert _count(s, "MUFU.TANH") == 64               # 2 hidden layers x 32 columns per thread
    assert _count(s, "FFMA2") >= 20 and _count(s, "FADD2") >= 12     # packed fp32x2 knot chains
    assert "F2FP.SATFINITE.F16" in s                  # fp16 operands, saturating conversion of the inputs
    # one spline element: 43 MUFU operations (profiles/fused_analysis_r01.md): 32 ex2 of the double softmaxes
    # + 2 x (ex2, lg2) of the double softplus + 6 rcp + 1 lg2
    # ... plus, statically, the cold side of the lazy exact-bin decision (one exact knot chain: 16 expf = 16
    # EX2 and the reciprocal refinements of its two softmaxes), taken by ~1e-4 of the elements
    mufu = _count(s, "MUFU.EX2") + _count(s, "MUFU.LG2") + _count(s, "MUFU.RCP")
    assert 43 <= mufu <= 82, mufu
    assert _count(s, "MUFU.EX2") <= 34 + 16 and _count(s, "MUFU.LG2") == 3


FUSED2 = "_ZN3nfk17nsf_fused2_kernelILi2ELb0ELb0ELb0ELb0EEEvNS_10Fused2ArgsE"        # FAST, forward, fp16 operands, production
FUSED2_SPLIT = "_ZN3nfk17nsf_fused2_kernelILi1ELb0ELb1ELb0ELb0EEEvNS_10Fused2ArgsE"  # HYBRID, forward, split operands


def test_second_generation_fused_kernel():
    """csrc/nsf_fused2.cu: warp-specialised roles with per-warpgroup register budgets (setmaxnreg), tcgen05 MMAs,
    TMA bulk ring, TMEM loads, packed fp32x2 bias adds / knot chains, 16-byte bias loads; launch allocation of
    80 registers and no spill in the production instantiations."""
    res = _usage()
    assert res[FUSED2]["REG"] == 80 and res[FUSED2]["STACK"] == 0 and res[FUSED2]["LOCAL"] == 0, res[FUSED2]
    s = _sass(FUSED2)
    assert _count(s, "USETMAXREG") == 3               # spline warps grow, hidden and MMA/TMA warpgroups shrink
    assert _count(s, "UTCHMMA") >= 14                 # GEMM1 (2) + GEMM2 piece (4) + one GEMM3 chunk (8), prologue copies
    assert _count(s, "UBLKCP") >= 4 and _count(s, "LDTM") >= 3
    assert _count(s, "MUFU.TANH") in (16, 32)         # hidden epilogue: 16 columns per step (two call sites)
    assert _count(s, "FADD2") >= 20 and _count(s, "FFMA2") >= 20
    assert _count(s, "LDS.128") >= 6                  # b3 as six 16-byte broadcast loads per element
    assert _count(s, "MUFU.LG2") == 3
    sp = _sass(FUSED2_SPLIT)
    assert _count(sp, "MUFU.TANH") == 0               # fp32-class tanh: ex2 + rcp
    assert _count(sp, "UTCHMMA") >= 3 * 14 - 4        # three MMAs per product (hi*hi + lo*hi + hi*lo)
    assert "F2FP" in sp and "HADD2.F32" in sp         # hi / lo operand split in the epilogues


def test_standalone_pairs_kernel_uses_tma_ring_and_packed_math():
    s = _sass(PAIRS)
    assert _count(s, "UBLKCP") >= 1 and _count(s, "SYNCS") >= 2       # TMA bulk loads completing on mbarriers
    assert _count(s, "FFMA2") >= 20
    assert _count(s, "SHFL") >= 5                     # log-det: warp-shuffle reduction (north star item 3)
    assert "LDG.E.64" in s or "LDG.E.EF.64" in s or re.search(r"LDG\.E\.[A-Z.]*64", s)      # 8-byte coalesced activations


def test_tensor_core_gemms_issue_tcgen05():
    res = _usage()
    for tag in ("gemm_ws_kernel", "wgrad_ws_kernel", "linear_tf32x3_kernel", "linear_bf16_kernel"):
        fn = next(k for k in res if tag in k)
        s = _sass(fn)
        assert _count(s, "UTCHMMA") >= 1, tag


BWD = "_ZN3nfk20nsf_fused_bwd_kernelILb0EEEvNS_12FusedBwdArgsE"


def test_layer_backward_kernel_instruction_mix_and_resources():
    """csrc/nsf_fused_bwd.cu: every product on the tensor cores, weights through the TMA ring, both knot chains of the
    adjoint in packed fp32x2, the conditioning-column adds as fire-and-forget reductions; 18 warps share four
    sub-partitions (five on two of them), so 96 registers is the ceiling, with at most a few spilled loop counters."""
    res = _usage()
    for k in (BWD, BWD.replace("ILb0E", "ILb1E")):
        assert res[k]["REG"] <= 96 and res[k]["STACK"] <= 64 and res[k]["LOCAL"] == 0, (k, res[k])
    s = _sass(BWD)
    # GEMM1 2 + GEMM2 4 + GEMM3 chunk 8 + dH2 K block 4 + dH1 4 + dXc 4 tcgen05.mma instructions in the issue loops
    assert _count(s, "UTCHMMA") == 26, _count(s, "UTCHMMA")
    assert _count(s, "UBLKCP") >= 2 and _count(s, "UBLKPF") == 2      # ring + W1 bulk copies, L2 prefetch of the next tile
    assert _count(s, "LDTM") >= 6                      # two tanh epilogues, chunk parameters (x16 + x8), two tanh backwards, dXc
    assert _count(s, "MUFU.TANH") == 64
    assert _count(s, "FFMA2") >= 30 and _count(s, "FMUL2") >= 30 and _count(s, "FADD2") >= 25
    assert _count(s, "REDG") == 8                      # conditioning-column gradient adds
    mufu = _count(s, "MUFU.EX2") + _count(s, "MUFU.LG2") + _count(s, "MUFU.RCP")
    assert 50 <= mufu <= 64, mufu                      # 32 ex2 of the double softmaxes + 4 x (ex2, lg2 | rcp) + the segment's divisions


def test_streaming_spline_kernel_runs_32_warps_per_sm():
    """csrc/rqs_coupling.cu, rqs_coupling_stream: 64 registers (eight 128-thread CTAs per SM), at most two spilled words,
    parameters through TMA bulk copies, no CTA-wide barrier inside the element loop (one after the prologue)."""
    res = _usage()
    ks = [k for k in res if "rqs_coupling_stream" in k]
    assert len(ks) == 6, ks                       # 3 arithmetics x 2 directions, K = 8 only
    for k in ks:
        assert res[k]["REG"] <= 64 and res[k]["STACK"] <= 16 and res[k]["LOCAL"] == 0, (k, res[k])
    s = _sass("_ZN3nfk19rqs_coupling_streamILi2ELi8ELb0EEEvNS_12CouplingArgsE")
    assert _count(s, "UBLKCP") >= 1
    assert _count(s, "BAR.SYNC") + _count(s, "BAR") <= 2, _count(s, "BAR")
