"""Steady-state time of MonteCarloEngine.run_point on the dense H_std graph for small block counts."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ldpc-simulator_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import load_code
from encoder_decoder_data import EncoderDecoderData
from mc_driver import MonteCarloEngine
name = sys.argv[1] if len(sys.argv) > 1 else "wimax_576_0.5"
edd = EncoderDecoderData(h=load_code(name).sparse_matrix())
eng = MonteCarloEngine(edd, graph="std", precision=sys.argv[2] if len(sys.argv) > 2 else "f64", max_iterations=20, seed=5)
for blocks in (50, 100, 200, 400, 1000, 4000):
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.run_point(3.0, 0.5, frames=blocks)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(blocks, "blocks", round(best * 1e3, 2), "ms", round(blocks / best), "frames/s", flush=True)
