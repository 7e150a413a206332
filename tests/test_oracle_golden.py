"""Pins the CPU oracle against the reference's own golden vectors:
the whole TestHivedAlgorithm scenario of pkg/algorithm/hived_algorithm_test.go replayed on the
oracle through the C ABI and the host-side mirror."""
from golden_scenario import Scenario


def test_oracle_reproduces_reference_golden_vectors(oracle_lib):
    sc = Scenario(oracle_lib)
    errors = sc.run()
    assert errors == [], "\n".join(errors)
    binds = [d for d in sc.decisions if d[1] == "bind"]
    assert len(binds) >= 25
