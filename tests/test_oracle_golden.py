"""CPU tier: the oracle (oracle/oracle.c) against the golden vectors frozen from the reference's own CUDA
extensions on a B200 (oracle/gen_golden.py -> tests/golden/ref_*.npz).  This is what pins the oracle."""
import pytest

from cases import CASES, compare, run_case
from conftest import golden_scales, load_golden


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    gold = load_golden(name)
    if gold is None:
        pytest.skip(f"tests/golden/ref_{name}.npz not generated yet (needs a GPU run of oracle/gen_golden.py)")
    from backends import OracleBackend
    got = run_case(name, OracleBackend(device_scales=golden_scales(gold)))
    # the oracle accumulates the MLP / grid gradients in fp32/fp64 while the reference uses fp16 accumulators
    # and atomics: same tolerances as the GPU parity tests
    compare(got, gold)


def test_cases_are_deterministic(oracle_backend):
    a = run_case("march_lego", oracle_backend)
    b = run_case("march_lego", oracle_backend)
    compare(a, b)
    assert int(a["counts"].sum()) == int(a["counter"][0]) and int(a["counter"][1]) == len(a["counts"])
    assert a["counts"].max() > 32 and (a["counts"] == 0).any(), "case must contain long rays and rays that miss"
