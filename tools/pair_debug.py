"""Debug: where do the pair kernel and the one-frame kernel start to differ?"""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ldpc-simulator_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_code
from settings import Settings
from spa_decoder import SPA_Decoder
import torch

class Edd:
    def __init__(self, h): self._h_sparse_cached, (self._m, self._n) = h, h.shape

code = load_code(sys.argv[1] if len(sys.argv) > 1 else "wimax_2304_0.5")
VARIANT = os.environ.get("PAIR_DEBUG_VARIANT", "")        # e.g. one_gather_kernel
def dec(it, one):
    st = Settings(); st.set_max_iterations(it); st.set_precision("f32_fast"); st.set_early_termination(False); st.set_one_frame_kernel(one)
    if VARIANT and not one:
        getattr(st, "set_" + VARIANT)(True)
    return SPA_Decoder(Edd(code.csr()), st)
rng = np.random.default_rng(3)
F = 2048
sig = 1 / np.sqrt(10 ** 0.2)
llr = torch.as_tensor((2 * (-1 + sig * rng.standard_normal((F, code.n))) / sig ** 2).astype(np.float32)).cuda()
for it in (1, 2, 3, 5, 20):
    a = dec(it, False).decode_batch_device(llr, want_posterior=True).post.cpu().numpy()
    a2 = dec(it, False).decode_batch_device(llr, want_posterior=True).post.cpu().numpy()
    b = dec(it, True).decode_batch_device(llr, want_posterior=True).post.cpu().numpy()
    b2 = dec(it, True).decode_batch_device(llr, want_posterior=True).post.cpu().numpy()
    d = a != b
    print(f"iters {it}: pair deterministic {np.array_equal(a, a2)}  one deterministic {np.array_equal(b, b2)}  differing entries {d.mean():.5f}"
          f"  frames with a difference {d.any(axis=1).mean():.4f}  max rel {np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3)):.3e}")
    if d.any() and it <= 2:
        f, j = np.argwhere(d)[0]
        cols = np.unique(np.argwhere(d)[:, 1] // 96)
        print("   first diff frame", f, "col", j, a[f, j], b[f, j], "column blocks with differences:", cols[:30], "frames parity", np.unique(np.argwhere(d)[:, 0] % 2))
