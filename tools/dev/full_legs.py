"""The two full-network side legs of bench.py (HandTrackNet train step at B=32 x N=4096, one tracked frame at B=1 x N=8192)
on our arm only -- for A/B runs with environment switches."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = bench.full_network_legs(torch, dev, "ours", 32, 4096, flush)
print("LEGS", {k: os.environ.get(k) for k in ("PN2_SEARCH_PREFETCH", "PN2_PDL", "PN2_KNN_COOP")},
      {k: (v.get("ms_per_step") or v.get("p50_ms")) for k, v in out.items() if isinstance(v, dict)})
