"""How close the deck-shaped 1000-step runs (tests/test_gpu_decks.py) are to their bars:
prints the largest deviations between device and oracle energies (run under gpurun)."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import psc_b200 as pb
import test_gpu_decks as t
from b200_helpers import gpu_state

for name in t.DECKS:
    d = t._deck(name)
    ref, _ = t._oracle_energies(name, 1000)
    for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        grid, mprts, mflds = gpu_state(d["og"], d["flds"], d["prts"], d["off"], dict(fma=0))
        psc = pb.Psc(grid, mflds, mprts, sort_interval=d["sort_interval"], marder_interval=d["marder_interval"],
                     marder_diffusion=0.9, marder_loop=1, fused=True)
        psc.initialize()
        got = [pb.api.energies(grid)]
        for s in range(1, 1001):
            psc.step()
            if s % 100 == 0:
                got.append(pb.api.energies(grid))
        got = np.array(got)
        grid.close()
        fg, fr = got[:, :6].sum(1), ref[:, :6].sum(1)
        print("%-22s field rel %.4f  field/total %.5f  comp/fieldmax %.4f  prt rel %.5f  total rel %.5f" % (
            name, np.abs(fg[1:] / fr[1:] - 1).max(), (np.abs(fg - fr) / ref.sum(1).min()).max(),
            np.abs(got[:, :6] - ref[:, :6]).max() / fr.max(), np.abs(got[:, 6:] / ref[:, 6:] - 1).max(),
            np.abs(got.sum(1) / ref.sum(1) - 1).max()))
