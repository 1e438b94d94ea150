"""CPU check (cuobjdump on the in-tree libpvb200.so) that the hot kernels are built from the hardware features DESIGN.md
claims for them -- the SASS mnemonics /opt/skills/guides/B200_PROFILING.md names as proof: UTCHMMA (tcgen05.mma), UBLKCP
(bulk async copy, TMA engine), SYNCS (mbarrier), LDGSTS (cp.async), FFMA2 (packed fp32 FMA), USETMAXREG (setmaxnreg) --
and that the FMA loops of the fp32 convolutions carry no scalar FFMA and no local-memory spills."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "predict_pv_yield_b200", "libpvb200.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    import __graft_entry__ as g

    g.build()
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out or "SM100a" in out.upper() or "arch = sm_100" in out, "library is not built for sm_100a"
    kernels = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            ins = line.split(None, 1)[1]
            ins = re.sub(r"^@!?U?P\w+\s+", "", ins)
            kernels[name].append(ins.split()[0].rstrip(";"))
    return kernels


def _find(kernels, *parts):
    hits = [k for k in kernels if all(p in k for p in parts)]
    assert hits, (parts, sorted(kernels)[:5])
    return hits


def _count(ops, prefix):
    return sum(1 for o in ops if o.split(".")[0] == prefix)


def test_bf16_convolutions_run_on_tcgen05_with_tma_operands(sass):
    for k in _find(sass, "conv3d_igemm_bf16_kernel") + _find(sass, "conv3d_wgrad_bf16_kernel") + _find(sass, "fc1_bf16_kernel"):
        ops = sass[k]
        assert _count(ops, "UTCHMMA") > 0, f"{k}: no tcgen05.mma"
        if "fc1_bf16_kernelILi2E" in k:  # fc1 weight gradient: both operands are activations, staged by cp.async
            assert _count(ops, "LDGSTS") > 0, f"{k}: no async copy"
        else:
            assert _count(ops, "UBLKCP") > 0, f"{k}: no bulk async copy"
        assert _count(ops, "SYNCS") > 0, f"{k}: no mbarrier"
        assert _count(ops, "HMMA") == 0, f"{k}: legacy mma.sync in a tcgen05 kernel"


def test_fp32_convolutions_run_on_packed_fma(sass):
    for k in _find(sass, "conv3d_direct_f32_kernel") + _find(sass, "conv3d_wgrad_f32"):
        ops = sass[k]
        n2 = _count(ops, "FFMA2")
        assert n2 >= 72, f"{k}: {n2} FFMA2"
        first = next(i for i, o in enumerate(ops) if o.startswith("FFMA2"))
        last = len(ops) - 1 - next(i for i, o in enumerate(reversed(ops)) if o.startswith("FFMA2"))
        loop = ops[first:last + 1]
        assert _count(loop, "FFMA") == 0, f"{k}: scalar FFMA inside the packed-FMA loop"
        assert _count(loop, "LDGSTS") == 0 or "ws" not in k, f"{k}: consumer loop of the warp-specialised kernel issues copies"


def test_warp_specialised_wgrad_uses_mbarriers_and_register_reallocation(sass):
    (wide,) = _find(sass, "conv3d_wgrad_f32_ws_wide_kernel")
    ops = sass[wide]
    assert _count(ops, "USETMAXREG") == 2, "setmaxnreg.dec (producers) + setmaxnreg.inc (consumers)"
    assert _count(ops, "SYNCS") > 0 and _count(ops, "LDGSTS") > 0
    first = next(i for i, o in enumerate(ops) if o.startswith("FFMA2"))
    last = len(ops) - 1 - next(i for i, o in enumerate(reversed(ops)) if o.startswith("FFMA2"))
    loop = ops[first:last + 1]
    assert _count(loop, "LDL") == 0 and _count(loop, "STL") == 0, "spills inside the FMA loop"
    # two unrolled iterations: 432 FFMA2 against ~24 shared-memory loads (the carried input window)
    assert _count(loop, "FFMA2") == 432 and _count(loop, "LDS") <= 30
    for k in _find(sass, "conv3d_wgrad_f32_ws_narrow_kernel") + _find(sass, "conv3d_wgrad_f32_ws_kernel"):
        assert _count(sass[k], "USETMAXREG") == 0


def test_fp32_mode_tensor_core_kernels(sass):
    """The fp32-mode convolutions of round 2: 3xTF32 implicit GEMM (single CTA and CTA pair) and the three-way bf16 split
    weight gradient -- tcgen05 MMAs, TMA-engine operands (bulk copies; tensor-map loads in the weight gradient), tensor
    memory loads in the epilogues, mbarriers; the pair kernel issues the 2-CTA form of the MMA and of the commit."""
    single = _find(sass, "conv3d_igemm_tf32x3_kernel")
    pair = _find(sass, "conv3d_igemm_tf32x3_pair_kernel")
    wgrads = _find(sass, "conv3d_wgrad_bf16x3_kernel")  # <3>: three-way bf16 split, <2>: two-way fp16 split
    assert len(wgrads) == 2
    for k in single + pair:
        ops = sass[k]
        ks = int(re.search(r"kernelILi(\d)E", k).group(1))
        assert _count(ops, "UTCHMMA") == 27 * ks, f"{k}: {_count(ops, 'UTCHMMA')} MMAs per plane, expected 27 x {ks}"
        assert _count(ops, "UBLKCP") > 0 and _count(ops, "LDTM") > 0 and _count(ops, "SYNCS") > 0, k
        two_cta = sum(1 for o in ops if o.startswith("UTCHMMA") and "2CTA" in o)
        assert two_cta == (27 * ks if k in pair else 0), f"{k}: {two_cta} 2-CTA MMAs"
    for wgrad in wgrads:
        ops = sass[wgrad]
        pieces = int(re.search(r"kernelILi(\d)E", wgrad).group(1))
        # per 16 positions and kw tap: six products of the three-way split, three of the two-way split (4 x 16 positions unrolled)
        assert _count(ops, "UTCHMMA") >= 4 * (6 if pieces == 3 else 3) and _count(ops, "UTMALDG") == 2 and _count(ops, "LDTM") >= 3, wgrad
        assert sum(1 for o in ops if o.startswith("F2FP")) >= 8 * pieces, "packed bf16 / fp16 conversion of the split warps"
        assert _count(ops, "F2F") == 0, "scalar F2F (quarter-rate pipe) in the split loop"


def test_fc1_tensor_core_and_row_step_kernels(sass):
    """fc1 of the fp32 head (three-way bf16 split): tcgen05 MMAs fed by tensor-map loads, the backward kernels leave through
    tensor-map STORES (UTMASTG) -- 4-byte register stores to rows 4.4 MB apart ran at a third of the speed; the bf16 row-step
    weight gradient: two tensor-map loads per step straight into the operand layout, no conversion instructions."""
    for name, stores in (("fc1x3_fwd_kernel", 0), ("fc1x3_dgrad_kernel", 1), ("fc1x3_wgrad_kernel", 1)):
        (k,) = _find(sass, name)
        ops = sass[k]
        assert _count(ops, "UTCHMMA") >= 12 and _count(ops, "LDTM") >= 1, k
        assert _count(ops, "UTMALDG") >= 1 and _count(ops, "UTMASTG") == stores, k  # (unrolled loops repeat the load)
        assert sum(1 for o in ops if o.startswith("F2FP")) >= 12 and _count(ops, "F2F") == 0, k
        assert _count(ops, "UTMAREDG") == (1 if name == "fc1x3_wgrad_kernel" else 0), k  # later batch chunks add to dW
    (k,) = _find(sass, "conv3d_wgrad_bf16_rows_kernel")
    ops = sass[k]
    assert _count(ops, "UTCHMMA") >= 12 and _count(ops, "UTMALDG") == 2 and _count(ops, "LDTM") >= 1, k
    assert sum(1 for o in ops if o.startswith("F2FP")) == 0, k
