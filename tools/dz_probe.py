"""dlogits pass: prefetch depth x L2 hints, ~1.5 s sustained each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hint_sweep import run  # noqa: E402
for st in (0, 5):
    os.environ["PRL_DZ_STAGES"] = str(st)
    for h in ((2, 1, 1), (2, 0, 1), (2, 0, 0), (0, 0, 0)):
        print(f"stages={st or 6}", end=" ")
        run("DZ", h, secs=1.5)
