"""Computes the config-5 reference scores with the CPU oracle (multi-threaded blocked NW, exact int64) and
freezes them in tests/golden/config5_scores.json.  ~15 min on 8 cores for the full 1 Mbp x 1 Mbp table."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomics_rs_b200 import workloads as wl
from oracle import gxo
a, b = wl.long_pair(1_000_000)
out = {"scores": dict(zip(("s_match", "s_mismatch", "g", "h"), wl.CONFIG_TOML)), "seeds": ["0x5EED1000", "0x5EED1001"], "prefix_scores": {}}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "config5_scores.json")
for n in [4096, 65536, 262144, 1_000_000]:
    t = time.time()
    sc = gxo.nw_score_blocked(a[:n], b[:n], wl.CONFIG_TOML, n_threads=8, blk=2048)
    out["prefix_scores"][str(n)] = sc
    print(n, sc, round(time.time() - t, 1), "s", flush=True)
    json.dump(out, open(path, "w"), indent=1)
