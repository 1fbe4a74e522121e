"""phylonium_b200.simgen against the oracle's simf restatement and the reference's simf."""
import oracle_lib
from phylonium_b200 import simgen


def test_simgen_equals_oracle_simf():
    o = oracle_lib.best()
    for seed, length, d in ((1, 1000, 0.01), (7, 5003, 0.1), (4, 70, 0.5), (11, 2000, 0.0), (2, 200000, 0.03)):
        assert simgen.simf(seed, seed + 3, length, d) == o.simf(seed, seed + 3, length, d)
    assert simgen.simf_set(3, 3000, [0.02, 0.05]) == oracle_lib.port().simf_set(3, 3000, [0.02, 0.05])
