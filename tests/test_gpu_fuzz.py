"""Seeded random cases from tools/fuzz_compositions.py and tools/fuzz_uservals.py (the tools take any seed and count; these
are the fixed ones of the suite): compositions of example filters that exist nowhere in the reference's tree, and example
filters away from their default arguments -- booleans pick the specialised kernel variants -- against the oracle."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.gpu
def test_random_compositions_match_oracle():
    import fuzz_compositions
    cases, failures = fuzz_compositions.run(2024, 30)
    assert cases == 30 and not failures, failures


@pytest.mark.gpu
@pytest.mark.parametrize("seed,count,only", [(7, 50, ""), (8, 12, "Droste")])
def test_random_arguments_match_oracle(seed, count, only):
    import fuzz_uservals
    cases, failures = fuzz_uservals.run(seed, count, only)
    assert cases == count and not failures, failures
