"""Where the deck tolerances of tests/test_gpu_decks.py come from.

North star: "field and particle energies over 1000 steps agree within 1 %".  Two of the four
deck-shaped runs have a field energy that is small and noise-driven (bubble: 2 % of the total
energy; flatfoil: no initial fields at all, 0.1 %), and a noise-driven quantity is chaotic:
the CPU ORACLE run twice on the same particles, once in the loader's order and once shuffled
inside every patch -- nothing but the summation order of J changes -- differs FROM ITSELF by
SELF_DEV in that quantity after 1000 steps.  No implementation whose J summation order differs
from the reference's (any GPU: atomics) can be held tighter than that, so the field energy of
those two decks is held to FIELD_FACTOR x SELF_DEV of itself AND to 1 % of the total energy;
everything else (particle energies, total energy, the field energy of the decks that start
with fields) to the contract's 1 %.

SELF_DEV is not folklore: tests/test_decks_chaos.py (CPU) re-measures it and fails if it
leaves [1/2, 2] x the value recorded here."""

# max over the 100-step samples of |E_field(shuffled) / E_field(original) - 1|, oracle vs oracle, 1000 steps
SELF_DEV = {"bubble_yz": 0.0161, "flatfoil_yz": 0.0263}
# ... and of the particle energies (sum over species), for scale: far inside the 1 % contract
SELF_DEV_PARTICLES = {"bubble_yz": 0.00026, "flatfoil_yz": 0.00094}
FIELD_FACTOR = {"bubble_yz": 5., "flatfoil_yz": 10.}


def field_rtol(name):
    """relative bound on the total field energy of deck `name`, device vs oracle"""
    if name in SELF_DEV:
        return max(1e-2, FIELD_FACTOR[name] * SELF_DEV[name])
    return 1e-2
