"""Build evidence (CPU, needs the CUDA toolkit's cuobjdump): the objects are sm_100a code and the hot kernels really use the
Blackwell tensor-core / TMA instructions the design claims -- tcgen05 MMAs (UTCHMMA, `.2CTA` for cta_group::2), TMA tile loads
(UTMALDG), tensor-memory loads (LDTM) -- and the peer-memory relay uses bulk copies (UBLKCP)."""
import os
import re
import shutil
import subprocess

import pytest

from vipant_b200 import build as vb_build

LIB_DIR = os.path.dirname(vb_build.LIB_PATH)
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")


def _sass(obj):
    path = os.path.join(LIB_DIR, obj)
    if not os.path.exists(path):
        pytest.skip(f"{obj} not built (python -m vipant_b200.build)")
    out = subprocess.run([CUOBJDUMP, "-sass", path], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out, f"{obj} holds no sm_100a code"
    return out


def _count(sass, pattern):
    return len(re.findall(pattern, sass))


def test_pair_kernels_are_tcgen05_cta_pairs():
    s = _sass("infonce_pair.o")
    assert _count(s, r"UTCHMMA\.2CTA") >= 16            # cta_group::2 MMAs of the three sweep kernels
    assert _count(s, r"UTMALDG\.2D\.2CTA") >= 8         # TMA tile loads accounted on the leader's barrier
    assert _count(s, r"LDTM") >= 3                      # accumulators read back from tensor memory (loops, not unrolled)
    assert _count(s, r"UBLKCP") >= 2                    # relay CTAs: cp.async.bulk peer -> shared -> local


def test_single_cta_sweep_is_tcgen05():
    s = _sass("infonce_tc.o")
    assert _count(s, r"UTCHMMA") >= 8 and _count(s, r"UTMALDG") >= 4 and _count(s, r"LDTM") >= 2


def test_encoder_tail_projection_is_a_tcgen05_cta_pair():
    s = _sass("encoder_tail.o")
    assert _count(s, r"UTCHMMA\.2CTA") >= 4 and _count(s, r"UTMALDG\.2D\.2CTA") >= 2 and _count(s, r"LDTM") >= 2


def test_no_legacy_tensor_core_instructions():
    """No mma.sync / wmma (HMMA) kernels anywhere in the library: the tensor-core work is tcgen05 only."""
    for obj in sorted(os.listdir(LIB_DIR)):
        if obj.endswith(".o"):
            s = _sass(obj)
            assert _count(s, r"\bHMMA\b") == 0, obj
