"""GPU parity of NeuralSplineCoupling.act (bijectors.py:319,345) other than the default swish: the conditioner
runs on the fp32 FFMA kernels (eval: chain_kernel; train: stored pre-activations + the generic GEMM family), checked
like the default path - eval passes against the oracle in fp32 / fp64, train gradients against float64 autograd."""
import numpy as np
import pytest

from oracle import torch_oracle as to
from oracle import zenflow_oracle as zo
from tests.helpers import assert_fp32_parity, product_chain, to64, trained_variables

pytestmark = pytest.mark.gpu

ACTS = ["relu", "tanh", "sigmoid", "gelu", "elu", "softplus", "leaky_relu"]
GRAD_RTOL = 1e-4


def _with_act(ops, act):
    return [dict(op, act=act) if op["kind"] == "coupling" else op for op in ops]


def _flow_vars(v):
    return {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}


@pytest.mark.parametrize("act", ACTS)
def test_eval_passes(act):
    """A chain the tensor-core kernel would take with swish (width 128, K = 16): the activation sends it to the
    FFMA kernel; forward, log-prob and inverse against the oracle."""
    from zenflow_b200 import Flow

    D, C, K, M = 3, 2, 16, 2311
    rng = np.random.default_rng(len(act))
    ops = _with_act(zo.make_chain(D, K, (128, 128)), act)
    x = rng.normal(0.3, 1.2, (M, D)).astype(np.float32)
    c = rng.uniform(0, 1, (M, C)).astype(np.float32)
    v = trained_variables(ops, x, c, seed=1)
    x64, c64 = x.astype(np.float64), c.astype(np.float64)
    chain = product_chain(ops)

    y, ld = chain.apply(v, x, c, train=False)
    yo, ldo, _ = zo.chain_forward(ops, v, x, c)
    y64, ld64, _ = zo.chain_forward(ops, to64(v), x64, c64)
    assert_fp32_parity(y, y64, yo, f"{act}: y", rtol=0, atol=5e-6, slack=6.0)
    assert_fp32_parity(ld, ld64, ldo, f"{act}: log_det", slack=3.0)

    flow = Flow(chain)
    lp = flow.apply(_flow_vars(v), x, c)
    lp64, _ = zo.flow_log_prob(ops, to64(v), x64, c64)
    lpo, _ = zo.flow_log_prob(ops, v, x, c)
    assert_fp32_parity(lp, lp64, lpo, f"{act}: log_prob", slack=3.0)

    u = np.random.default_rng(3).beta(12, 12, (M, D)).astype(np.float32)
    xi = chain.apply(v, u, c, method="inverse")
    xi64 = zo.chain_inverse(ops, to64(v), u.astype(np.float64), c64)
    xio = zo.chain_inverse(ops, v, u, c)
    assert_fp32_parity(xi, xi64, xio, f"{act}: inverse", rtol=0, atol=5e-6 * max(1.0, np.abs(xi64).max()), slack=3.0)

    # the default activation on the same variables gives a different answer (the field is not ignored)
    lp_swish = Flow(product_chain(_with_act(ops, "swish"))).apply(_flow_vars(v), x, c)
    assert np.abs(lp_swish - lp).max() > 1e-3


# smooth activations: float64 autograd is a stable truth; relu / leaky_relu have a kink at 0, where a pre-activation
# that rounds to the other side in fp32 flips a whole unit's contribution - they are covered by the eval test above
# and by the step-level check below
@pytest.mark.parametrize("act", ["tanh", "sigmoid", "gelu", "elu", "softplus"])
@pytest.mark.parametrize("shape", [(4, 2, 8, (16, 16), 300), (2, 1, 16, (128, 128), 515)], ids=["D4K8", "D2K16w128"])
def test_train_step_gradients_match_autograd(act, shape):
    from zenflow_b200 import Flow
    from zenflow_b200._train import TrainEngine

    D, C, K, layers, M = shape
    rng = np.random.default_rng(M)
    ops = _with_act(zo.make_chain(D, K, layers), act)
    x = rng.normal(0.3, 1.0, (M, D)).astype(np.float32)
    c = rng.uniform(0, 1, (M, C)).astype(np.float32)
    v = zo.init_variables(ops, D, C, 2, weight_scale=1.5, randomize_bn=True)
    loss64, g64, st64, gc64, lp64 = to.loss_and_grads(ops, to64(v), x.astype(np.float64), c.astype(np.float64))
    flow = Flow(product_chain(ops))
    flow.latent._latch_dim(D)
    eng = TrainEngine(flow, _flow_vars(v), D, C, micro_batch=128)
    lp_sum, gc_dev = eng.step(x, c, update=False, want_gc=True)
    loss = -float(lp_sum.item()) / M
    assert abs(loss - loss64) <= 1e-5 * abs(loss64) + 1e-5
    grads = eng.gradients()["bijector"]
    for name, layers_ in g64.items():
        for lname, leaves in layers_.items():
            for leaf, ref in leaves.items():
                got = grads[name][lname][leaf].cpu().numpy()
                scale = np.abs(ref).max() + 1e-12
                e = np.abs(got - ref).max() / scale
                assert e <= GRAD_RTOL, f"{act}: {name}/{lname}/{leaf}: rel err {e:.2e} (scale {scale:.2e})"
    gc = gc_dev.cpu().numpy()
    assert np.abs(gc - gc64).max() <= GRAD_RTOL * np.abs(gc64).max()


@pytest.mark.parametrize("act", ["relu", "leaky_relu"])
def test_train_step_gradients_piecewise_linear(act):
    """relu / leaky_relu: the same check with the tolerance a handful of flipped units allows (the loss itself is
    continuous across the kink and is held to the usual 1e-5)."""
    from zenflow_b200 import Flow
    from zenflow_b200._train import TrainEngine

    D, C, K, layers, M = 4, 2, 8, (16, 16), 300
    rng = np.random.default_rng(M)
    ops = _with_act(zo.make_chain(D, K, layers), act)
    x = rng.normal(0.3, 1.0, (M, D)).astype(np.float32)
    c = rng.uniform(0, 1, (M, C)).astype(np.float32)
    v = zo.init_variables(ops, D, C, 2, weight_scale=1.5, randomize_bn=True)
    loss64, g64, st64, gc64, lp64 = to.loss_and_grads(ops, to64(v), x.astype(np.float64), c.astype(np.float64))
    flow = Flow(product_chain(ops))
    flow.latent._latch_dim(D)
    eng = TrainEngine(flow, _flow_vars(v), D, C, micro_batch=128)
    lp_sum, _ = eng.step(x, c, update=False, want_gc=True)
    loss = -float(lp_sum.item()) / M
    assert abs(loss - loss64) <= 1e-5 * abs(loss64) + 1e-5
    grads = eng.gradients()["bijector"]
    for name, layers_ in g64.items():
        for lname, leaves in layers_.items():
            for leaf, ref in leaves.items():
                got = grads[name][lname][leaf].cpu().numpy()
                e = np.abs(got - ref).max() / (np.abs(ref).max() + 1e-12)
                assert e <= 2e-3, f"{act}: {name}/{lname}/{leaf}: rel err {e:.2e}"
