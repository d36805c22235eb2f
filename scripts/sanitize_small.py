"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import zenflow_oracle as zo
from tests.helpers import product_chain, trained_variables
from zenflow_b200 import Flow
from zenflow_b200.distributions import Beta
from zenflow_b200._train import TrainEngine
from zenflow_b200.utils import rqs_forward_raw, rqs_inverse_raw

rng = np.random.default_rng(0)
for (M, d, K) in [(1031, 1, 16), (517, 8, 32), (300, 3, 5)]:
    th = rng.standard_normal((M, d, 3 * K - 1)).astype(np.float32); x = rng.uniform(-.1, 1.1, (M, d)).astype(np.float32)
    rqs_forward_raw(x, th, K); rqs_inverse_raw(x, th, K)
for impl in ("", "simt", "umma8"):
    if impl: os.environ["ZF_CHAIN_IMPL"] = impl
    else: os.environ.pop("ZF_CHAIN_IMPL", None)
    for (D, C, K, layers, nc, roll, M) in [(2, 1, 16, (128, 128), None, 1, 777), (16, 4, 32, (128, 128), 3, 2, 300), (5, 3, 7, (64, 48), None, 1, 200),
                                           (24, 8, 16, (128, 128), 2, 3, 390), (6, 1, 32, (128, 128), 2, 1, 270)]:
        ops = zo.make_chain(D, K, layers, n_couplings=nc, roll_shift=roll)
        x = rng.normal(0.3, 1.1, (M, D)).astype(np.float32); c = rng.uniform(0, 1, (M, C)).astype(np.float32) if C else None
        v = trained_variables(ops, x, c)
        flow = Flow(product_chain(ops), latent=Beta()); flow.latent._latch_dim(D)  # fresh latent: the default one is shared
        fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
        flow.apply(fv, x, c); flow.apply(fv, x, c, method="inverse")
        flow.apply(fv, c if C else M, method="sample")
os.environ.pop("ZF_CHAIN_IMPL", None)
for gemm in ("", "simt"):
    if gemm: os.environ["ZF_GEMM_IMPL"] = gemm
    for (D, C, K, layers, nc, roll, M) in [(16, 4, 32, (128, 128), 2, 2, 700), (3, 0, 5, (40,), None, 1, 333), (24, 8, 16, (128, 128), 2, 3, 390),
                                           (6, 1, 32, (128, 128), 2, 1, 270)]:
        ops = zo.make_chain(D, K, layers, n_couplings=nc, roll_shift=roll)
        x = rng.normal(0.3, 1.0, (M, D)).astype(np.float32); c = rng.uniform(0, 1, (M, C)).astype(np.float32) if C else None
        v = zo.init_variables(ops, D, C, 2, randomize_bn=True)
        flow = Flow(product_chain(ops), latent=Beta()); flow.latent._latch_dim(D)  # fresh latent: the default one is shared
        eng = TrainEngine(flow, {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}, D, C, micro_batch=256)
        eng.step(x, c); eng.step(x, c)
torch.cuda.synchronize()
print("sanitize run complete")
