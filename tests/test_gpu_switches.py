"""The experiment switches of libb2s are read once per process, so each variant runs tools/experiments/switch_check.py
in its own interpreter: every alternative kernel that ships in the library (measured slower or kept for comparison,
DESIGN.md 4.1-4.3) is executed and checked on the GPU, not just compiled."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = [
    {},                                              # defaults
    {"B2S_TC_SPLIT_FOLD": "1"},                      # offset-split reduction folded into the convolution kernel
    {"B2S_TC_ASYNC": "1"},                           # cp.async gather ring
    {"B2S_TC_CTAS": "3"},                            # 3 CTAs / SM entry point
    {"B2S_TC_PERSIST": "0"},                         # round-1 per-tile tcgen05 kernel
    {"B2S_CONV_WS": "1"},                            # warp-stream mma.sync convolution for the narrow layers
    {"B2S_BN_FUSED": "0"},                           # two-kernel BatchNorm
    {"B2S_WGRAD_DET": "0"},                          # shared-memory mma.sync weight gradient (atomics)
    {"B2S_WGRAD_DET": "0", "B2S_WGRAD_MMA": "0"},    # fp32 FMA weight gradient
]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()) or "defaults")
def test_switch_variant(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "experiments", "switch_check.py")], env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SWITCH_CHECK_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
