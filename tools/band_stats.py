"""GX_FILL_STATS=1 python tools/band_stats.py [length] [bands] -- wait breakdown of the fill kernel on the config-5 pair"""
import os, sys
os.environ["GX_FILL_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib, workloads as wl
_lib.ensure_init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
bands = int(sys.argv[2]) if len(sys.argv) > 2 else 1
a, b = wl.long_pair(n)
band = gx.Band(n, n, bands, 0, bands, wl.CONFIG_TOML)
band.upload(a, b)
for _ in range(2):
    band.execute()
top, bnd, tile, s1, nt = [band.stat(k) for k in range(10, 15)]
K = int(band.stat(15))
steps = nt * 4096 + 0.0
print(f"nw {n}x{n} bands={bands}: K={K} chain1={int(band.stat(17))} fill {band.fill_ms:.3f} ms = {(n+1)*(n+1)/band.fill_ms/1e6:.0f} GCUPS, tiles {int(nt)}")
print(f"  warp cycles inside tiles {tile:.3e}: top wait {100*top/tile:.1f}%  left-boundary wait {100*bnd/tile:.1f}%  s1 TMA wait {100*s1/tile:.1f}%")
print(f"  avg tile {tile/nt/1.965e3:.1f} us -> {(tile-top-bnd-s1)/nt/min(4096,n):.0f} clk per step excluding waits, {tile/nt/min(4096,n):.0f} including")
