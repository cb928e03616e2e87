"""GPU parity tests of the candidate-search kernel through the C ABI against the oracle."""
import numpy as np
import pytest

import cs_cases
from oracle_lib import CsOracle

pytestmark = pytest.mark.gpu


def test_cs_search_matches_oracle(aligner):
    from ngmlr_b200 import refindex
    contigs = cs_cases.genome_contigs()
    orc = CsOracle([c.tobytes() for c in contigs])
    idx = refindex.build_index(refindex.encode_reference(contigs))
    aligner.set_index(idx)
    subs = cs_cases.subreads(400, 23, contigs)
    got, mx = aligner.cs_search(subs)
    for i, s in enumerate(subs):
        want, wmx = orc.search(s)
        assert got[i] == want, (i, len(s), got[i][:3], want[:3])
        assert mx[i] == np.float32(wmx)
    # a second batch on the same context, different sensitivity
    got2, _ = aligner.cs_search(subs[:50], sensitivity=0.5)
    for i in range(50):
        assert got2[i] == orc.search(subs[i], sensitivity=0.5)[0]
    orc.close()
