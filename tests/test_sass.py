"""The built library is Blackwell-native: its tensor-core work is tcgen05 (UTCHMMA with TMEM loads / stores and TMA),
never the warp-level mma.sync / wgmma paths.  CPU test: disassembles pytorch_glow_b200/libglowk.so with cuobjdump."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch_glow_b200", "libglowk.so")


@pytest.mark.skipif(shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"),
                    reason="cuobjdump not available")
def test_tensor_core_kernels_are_tcgen05():
    import __graft_entry__ as g
    g.build()
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out
    ops = collections.Counter()
    per_fn = collections.defaultdict(collections.Counter)
    fn = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops[m.group(1)] += 1
            per_fn[fn][m.group(1)] += 1
    assert ops["UTCHMMA"] > 100 and ops["LDTM"] > 0 and ops["UTMALDG"] > 0 and ops["UTMASTG"] > 0
    assert ops["HMMA"] == 0 and ops["HGMMA"] == 0 and ops["IMMA"] == 0
    # every instance of the fused coupling-net kernel writes an operand back to tensor memory (tcgen05.st) and issues
    # TS-mode MMAs; the bit-mask backward instances stage their masks with cp.async (LDGSTS)
    fused = {k: v for k, v in per_fn.items() if k and "cnet_chain_kernel" in k}
    assert len(fused) == 8
    for name, c in fused.items():
        assert c["STTM"] >= 1 and c["UTCHMMA"] >= 32 and c["LDTM"] >= 4, name
    assert sum(1 for c in fused.values() if c["LDGSTS"] > 0) == 2
