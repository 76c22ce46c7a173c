"""not gpu: the built library is Blackwell-native where the north star says so.  `cuobjdump -sass` of libasn_b200.so must
show tcgen05.mma (UTCHMMA, incl. the cta_group::2 form), TMEM loads (LDTM), TMA loads and stores (UTMALDG / UTMASTG),
tcgen05.commit (UTCBAR) and mbarriers (SYNCS) -- and no legacy HMMA (mma.sync / wmma) anywhere.  The per-kernel table is
profiles/r02_sass_histogram.md (tools/sass_histogram.py)."""
import os
import re
import shutil
import subprocess
from collections import Counter

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adaptsegnet_b200", "lib", "libasn_b200.so")


def _sass_by_kernel():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    return kernels


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_library_sass_is_blackwell_native():
    from adaptsegnet_b200 import _build

    _build.build()
    kernels = _sass_by_kernel()
    assert len(kernels) > 40, len(kernels)
    total = Counter()
    for c in kernels.values():
        total.update(c)

    def count(prefix):
        return sum(n for op, n in total.items() if op.startswith(prefix))

    assert count("UTCHMMA") > 100 and any(".2CTA" in op for op in total if op.startswith("UTCHMMA"))
    assert count("LDTM") > 0 and count("UTMALDG") > 0 and count("UTMASTG") > 0 and count("UTCBAR") > 0 and count("SYNCS") > 0
    assert count("HMMA") == 0, "legacy mma.sync / wmma tensor-core instructions in the library"
    # every tensor-core kernel of the path issues tcgen05.mma: the persistent ring kernel (all modes) and the halo-tile kernels
    tc = [k for k, c in kernels.items() if any(op.startswith("UTCHMMA") for op in c)]
    assert any("umma_kernel" in k for k in tc) and sum("halo" in k for k in tc) >= 3, tc
    # the layout kernels of round 2 are in the binary (256-bit stores in the dYcol kernel)
    dycol = [c for k, c in kernels.items() if "aspp_dycols_chunk_kernel" in k]
    assert dycol and any(op.startswith("STG") and ".256" in op for op in dycol[0]), dycol
