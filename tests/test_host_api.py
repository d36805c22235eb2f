"""CPU checks of the host-side mirror of the FLAX-module API: variable-tree naming, init,
error behaviour (the things the reference's tests pin about module structure)."""
import numpy as np
import pytest

from zenflow_b200 import Flow
from zenflow_b200 import bijectors as bi
from zenflow_b200 import distributions as dist
from zenflow_b200.module import Module


def _shapes(t):
    return {k: _shapes(v) if isinstance(v, dict) else tuple(v.shape) for k, v in t.items()}


def test_bijector_is_abstract():
    """tests/test_bijectors.py:12-31."""
    with pytest.raises(TypeError):
        bi.Bijector()

    class Foo(bi.Bijector):
        def __call__(self, x, c, train=False):
            return super().__call__(x, c, train)

        def inverse(self, x, c):
            return super().inverse(x, c)

    foo = Foo()
    with pytest.raises(NotImplementedError):
        foo(1, 2)
    with pytest.raises(NotImplementedError):
        foo.inverse(1, 2)


def test_flow_variable_tree_matches_flax_layout():
    """SURVEY §8b; examples/deep_set.ipynb:466-486 shows the reference tree."""
    f = Flow(bi.rolling_spline_coupling(2))
    v = f.init(0, np.zeros((1, 2), np.float32), np.zeros((1, 1), np.float32))
    s = _shapes(v)
    assert set(s) == {"params", "batch_stats"}
    assert set(s["params"]["bijector"]) == {"bijectors_1", "bijectors_3"}
    assert s["params"]["bijector"]["bijectors_1"] == {
        "BatchNorm_0": {"scale": (2,), "bias": (2,)},
        "Dense_0": {"kernel": (2, 128), "bias": (128,)},
        "Dense_1": {"kernel": (128, 128), "bias": (128,)},
        "Dense_2": {"kernel": (128, 47), "bias": (47,)},
    }
    bs = s["batch_stats"]["bijector"]
    assert bs["bijectors_0"] == {"xmin_0": (1,), "xmax_0": (1,), "xmin_1": (1,), "xmax_1": (1,)}
    assert bs["bijectors_1"] == {"BatchNorm_0": {"mean": (2,), "var": (2,)}}
    st = v["batch_stats"]["bijector"]["bijectors_0"]
    assert np.isposinf(st["xmin_0"]).all() and np.isneginf(st["xmax_1"]).all()  # bijectors.py:243-248
    bn = v["params"]["bijector"]["bijectors_1"]["BatchNorm_0"]
    assert (bn["scale"] == 1).all() and (bn["bias"] == 0).all()
    k = v["params"]["bijector"]["bijectors_1"]["Dense_1"]["kernel"]
    assert abs(k.std() - 1 / np.sqrt(128)) < 0.01 and np.abs(k).max() <= 2 / np.sqrt(128) / 0.8796 + 1e-6
    assert f.latent.dim == 2  # lazily latched by the first evaluation (distributions.py:31-32)


def test_standalone_and_chain_naming():
    """tests/test_bijectors.py:43-46,200-202."""
    sb = bi.ShiftBounds(margin=0.01)
    v = sb.init(0, np.array([[1, 5], [3, 4], [6, 2]]), None)
    assert set(v) == {"batch_stats"} and set(v["batch_stats"]) == {"xmin_0", "xmax_0", "xmin_1", "xmax_1"}
    ch = bi.Chain([bi.ShiftBounds(margin=0.0), bi.Roll()])
    v = ch.init(0, np.zeros((3, 3), np.float32), None)
    assert set(v["batch_stats"]) == {"bijectors_0"}
    assert len(ch) == 2 and isinstance(ch[1], bi.Roll) and len(ch[0:1]) == 1
    assert bi.Roll().init(0, np.zeros((3, 2)), None) == {}


def test_bounded_columns_keep_no_statistics_and_validation():
    sb = bi.ShiftBounds(bounds=[(0, -1, 1), (1, 10, None), (2, None, 1)])
    v = sb.init(0, np.zeros((4, 3), np.float32), None)
    assert set(v["batch_stats"]) == {"xmin_1", "xmax_1", "xmin_2", "xmax_2"}
    with pytest.raises(ValueError):  # bijectors.py:169-171
        bi.ShiftBounds(bounds=[(5, 0, 1)]).init(0, np.zeros((4, 3), np.float32), None)
    with pytest.raises(ValueError):  # bijectors.py:172-174
        bi.ShiftBounds(bounds=[(0, 1, 0)]).init(0, np.zeros((4, 3), np.float32), None)
    with pytest.raises(ValueError):  # bijectors.py:156-158
        bi.ShiftBounds(margin=-0.1).init(0, np.zeros((4, 3), np.float32), None)
    with pytest.raises(ValueError):  # bijectors.py:159-161
        bi.ShiftBounds(margin=1.0).init(0, np.zeros((4, 3), np.float32), None)


def test_split_and_factory():
    """tests/test_bijectors.py:238-242, 259-266."""
    x = np.zeros((3, 3))
    xt, xc = bi.NeuralSplineCoupling._split(x)
    assert xt.shape[1] == 1 and xc.shape[1] == 2
    rsc = bi.rolling_spline_coupling(3, layers=(64, 64))
    kinds = [type(b).__name__ for b in rsc]
    assert kinds == ["ShiftBounds", "NeuralSplineCoupling", "Roll", "NeuralSplineCoupling", "Roll",
                     "NeuralSplineCoupling"]
    with pytest.raises(ValueError):
        bi.rolling_spline_coupling(0)
    with pytest.raises(ValueError):
        bi.rolling_spline_coupling(1)
    pre = bi.rolling_spline_coupling(2, preprocessing=[bi.Roll()])
    assert type(pre[0]).__name__ == "Roll"
    with pytest.raises(AssertionError):  # bijectors.py:326
        bi.NeuralSplineCoupling().init(0, np.zeros((2, 1), np.float32), None)


def test_distributions_repr_and_validation():
    """tests/test_distributions.py:25,83-86."""
    assert repr(dist.Uniform()) == "Uniform()"
    assert repr(dist.Beta()) == "Beta(peakness=12.0)"
    with pytest.raises(ValueError):
        dist.Beta(-1)
    assert dist.Normal().dim is None


def test_unbound_module_and_immutable_collections():
    sb = bi.ShiftBounds()
    with pytest.raises(RuntimeError):
        sb(np.zeros((2, 2), np.float32))
    assert isinstance(sb, Module)
    f = Flow(bi.ShiftBounds())
    v = f.init(0, np.zeros((3, 2), np.float32))
    assert set(v) == {"batch_stats"}  # no params collection at all (tests/test_flow.py:9)
    with pytest.raises(ValueError):
        f.apply(v, np.zeros((3, 2), np.float32), method="_steps_not_chain") if False else f.apply(
            v, np.zeros((3, 2), np.float32), method="_steps")
