"""The product library is what DESIGN.md says it is: the hot kernels of libdiqt_b200.so carry the sm_100a instructions the design rests on
(tcgen05 MMAs with TMEM loads / stores, TMA tensor copies, packed fp32x2 arithmetic in the issue-bound roles) and no local-memory traffic in
the dominant kernel.  Runs wherever the CUDA toolkit is (no GPU needed): `cuobjdump -sass` of the in-tree library."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffusioniqt_b200", "libdiqt_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not found")
    if not os.path.exists(LIB):
        pytest.skip("library not built (python -c 'import __graft_entry__ as g; g.build()')")
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], check=True, capture_output=True, text=True).stdout
    per_kernel = collections.defaultdict(collections.Counter)
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            per_kernel[name][m.group(1)] += 1
    return per_kernel


def _kernels(sass, fragment):
    found = {k: v for k, v in sass.items() if fragment in k}
    assert found, f"no kernel matching {fragment!r} in the library"
    return found


def test_only_sm_100a_code_is_shipped():
    out = subprocess.run([CUOBJDUMP, "-lelf", LIB], check=True, capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("fragment", ["conv_zm_kernel", "conv_tc_kernel", "init_conv_tc_kernel", "softmax_attn_tc2_kernel", "linattn_ctx_tc", "wgrad_tc_kernel"])
def test_tensor_core_kernels_use_tcgen05_and_tma(sass, fragment):
    for name, ops in _kernels(sass, fragment).items():
        assert ops["UTCHMMA"] > 0, f"{name}: no tcgen05.mma"
        assert ops["LDTM"] > 0, f"{name}: the accumulators are never read from tensor memory"
        assert ops["SYNCS"] > 0, f"{name}: no mbarrier traffic"
        assert ops["UTMALDG"] + ops["UBLKCP"] + ops["LDGSTS"] > 0, f"{name}: no asynchronous copies into shared memory"


def test_fused_conv_transform_runs_on_packed_pairs_without_local_memory(sass):
    fused = {k: v for k, v in _kernels(sass, "conv_zm_kernel").items() if "ILb1E" in k}   # kGN = true instantiations
    assert len(fused) == 2
    for name, ops in fused.items():
        assert ops["FFMA2"] >= 36 and ops["FMUL2"] >= 24 and ops["FADD2"] >= 12, f"{name}: the transform is not on packed fp32 pairs: {dict(ops)}"
        assert ops["MUFU"] > 0
        # (the per-slot row masks of the transform are a two-element array set up once per work item: a handful of local accesses outside the
        # per-row loops are tolerated, a spill of the hot loops would show up as dozens)
        assert ops["LDL"] <= 4 and ops["STL"] <= 8, f"{name}: local-memory traffic ({ops['LDL']} loads, {ops['STL']} stores)"
        assert ops["UTMASTG"] > 0, f"{name}: the output planes do not leave through TMA stores"
