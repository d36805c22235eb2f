"""Host-side sequencing of the train-mode phases (C ABI: "train step phases").

Two users:
  * the bijector / Flow modules when called with ``train=True`` (``apply(..., train=True,
    mutable=["batch_stats"])`` as in train.py:66-72 and the reference's tests);
  * ``TrainEngine``: the whole ``step`` of train.py:80-86 (loss, gradients, optimiser update)
    on flat parameter / gradient / moment buffers, optionally data-parallel: each rank holds a
    shard of the batch and the small batch statistics (ShiftBounds min/max, BatchNorm moments
    forward and backward) and the flat gradient are all-reduced with NCCL through
    ``torch.distributed`` so the result equals the single-device step on the whole batch.

No arithmetic happens in Python here: torch only owns the buffers, views and collectives.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._device import like_input, ptr, require_cuda, stream_ptr, to_device_f32

BN_MOMENTUM = 0.99  # flax.linen.BatchNorm default


# ---------------------------------------------------------------------------------------------
# small helpers
# ---------------------------------------------------------------------------------------------
def _dist_world(group):
    import torch.distributed as dist

    if group is None and not (dist.is_available() and dist.is_initialized()):
        return None, 1
    return dist, dist.get_world_size(group)


def _allreduce(t: torch.Tensor, group, op: str = "sum"):
    dist, world = _dist_world(group)
    if world == 1:
        return
    opmap = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}
    dist.all_reduce(t, op=opmap[op], group=group)


class DpComm:
    """The NCCL communicators of the data-parallel train step, created through the C ABI (zf_dp_*).

    Two communicators over the same ranks: one for the small batch statistics on the compute stream and one for
    the gradient buckets on a side stream (so a bucket's all-reduce overlaps the next coupling's backward).
    torch.distributed is only the rendezvous: it broadcasts the two ncclUniqueIds (any backend)."""

    def __init__(self, group=None):
        dist, world = _dist_world(group)
        self.world = world
        self.rank = dist.get_rank(group) if world > 1 else 0
        self.stats = None
        self.grads = None
        self.aux_stream = None
        if world == 1:
            return
        lib = _lib.load()
        src = dist.get_global_rank(group, 0) if group is not None else 0
        handles = []
        for _ in range(2):
            buf = (C.c_ubyte * _lib.DP_UNIQUE_ID_BYTES)()
            if self.rank == 0:
                _lib.check(lib.zf_dp_unique_id(buf), "zf_dp_unique_id")
            obj = [bytes(buf)]
            dist.broadcast_object_list(obj, src=src, group=group)
            h = C.c_void_p()
            idbuf = (C.c_ubyte * _lib.DP_UNIQUE_ID_BYTES).from_buffer_copy(obj[0])
            _lib.check(lib.zf_dp_comm_create(idbuf, self.rank, world, C.byref(h)), "zf_dp_comm_create")
            handles.append(h)
        self.stats, self.grads = handles
        self.aux_stream = torch.cuda.Stream()

    def close(self):
        lib = _lib.load()
        for h in (self.stats, self.grads):
            if h is not None:
                lib.zf_dp_comm_destroy(h)
        self.stats = self.grads = None


def _sb_struct(kinds, lo, hi, margin, xmin: torch.Tensor, xmax: torch.Tensor) -> _lib.ZfShiftBounds:
    sb = _lib.ZfShiftBounds()
    for i, k in enumerate(kinds):
        sb.kind[i] = int(k)
        sb.lo[i] = float(lo[i])
        sb.hi[i] = float(hi[i])
    sb.margin = float(margin)
    sb.xmin = ptr(xmin)
    sb.xmax = ptr(xmax)
    return sb


def _chain_struct(dim: int, cdim: int, ops: Sequence[_lib.ZfOp]):
    arr = (_lib.ZfOp * max(len(ops), 1))(*ops)
    ch = _lib.ZfChain()
    ch.dim, ch.cdim, ch.n_ops = dim, cdim, len(ops)
    ch.ops = C.cast(arr, C.POINTER(_lib.ZfOp))
    return ch, arr


def _op_sb(sb):
    op = _lib.ZfOp()
    op.kind = _lib.OP_SHIFT_BOUNDS
    op.shift_bounds = C.pointer(sb)
    return op


def _op_cp(cp):
    op = _lib.ZfOp()
    op.kind = _lib.OP_COUPLING
    op.coupling = C.pointer(cp)
    return op


def _op_roll(shift):
    op = _lib.ZfOp()
    op.kind = _lib.OP_ROLL
    op.shift = int(shift)
    return op


def _coupling_struct(mod, scale, bias, mean, var, kernels, biases) -> _lib.ZfCoupling:
    """zf_coupling of the NeuralSplineCoupling module `mod` over the given leaves."""
    hidden = mod.layers
    cp = _lib.ZfCoupling()
    cp.knots = int(mod.knots)
    cp.n_hidden = len(hidden)
    cp.act = int(getattr(mod, "_act_kind", 0))
    for i, w in enumerate(hidden):
        cp.hidden[i] = int(w)
    cp.bn_scale, cp.bn_bias, cp.bn_mean, cp.bn_var = ptr(scale), ptr(bias), ptr(mean), ptr(var)
    for i, (k, b) in enumerate(zip(kernels, biases)):
        cp.kernel[i] = ptr(k)
        cp.bias[i] = ptr(b)
    return cp


def _run_chain_forward(ch, x, c, y, ld, acc: bool):
    lib = _lib.load()
    M = x.shape[0]
    nbytes = int(lib.zf_chain_workspace_bytes(C.byref(ch), M))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
    fn = lib.zf_chain_forward_acc if acc else lib.zf_chain_forward
    _lib.check(fn(stream_ptr(), C.byref(ch), ptr(x), ptr(c), M, ptr(y), ptr(ld), ptr(ws), nbytes), "zf_chain_forward")


def _minmax_update(sb, x, D, group=None):
    lib = _lib.load()
    mm = torch.empty(2 * D, dtype=torch.float32, device=x.device)
    scratch = torch.empty(2 * D, dtype=torch.int32, device=x.device)
    _lib.check(lib.zf_shift_bounds_minmax(stream_ptr(), C.byref(sb), ptr(x), x.shape[0], D, ptr(mm), ptr(scratch)),
               "zf_shift_bounds_minmax")
    dist, world = _dist_world(group)
    if world > 1:   # one all-reduce(min) over [min | -max]
        mm[D:].neg_()
        _allreduce(mm, group, "min")
        mm[D:].neg_()
    _lib.check(lib.zf_shift_bounds_update(stream_ptr(), C.byref(sb), D, ptr(mm)), "zf_shift_bounds_update")


def _bn_batch_stats(x, c, D, Cdim, global_count, bmean, bvar, ra_mean, ra_var, group=None):
    lib = _lib.load()
    F = D - D // 2 + Cdim
    sums = torch.empty(2 * F, dtype=torch.float64, device=x.device)
    _lib.check(lib.zf_bn_moments(stream_ptr(), ptr(x), ptr(c), x.shape[0], D, Cdim, ptr(sums)), "zf_bn_moments")
    _allreduce(sums, group, "sum")
    _lib.check(lib.zf_bn_finalize(stream_ptr(), ptr(sums), float(global_count), F, BN_MOMENTUM, ptr(bmean), ptr(bvar),
                                  ptr(ra_mean), ptr(ra_var)), "zf_bn_finalize")


# ---------------------------------------------------------------------------------------------
# module-level train-mode calls (apply(..., train=True, mutable=["batch_stats"]))
# ---------------------------------------------------------------------------------------------
def _leaf_out(t: torch.Tensor, template):
    return t if isinstance(template, torch.Tensor) else t.cpu().numpy()


def shift_bounds_train(mod, x):
    """ShiftBounds.__call__(train=True), bijectors.py:164-208,250-260."""
    dev = require_cuda()
    xd = to_device_f32(x, dev)
    D = xd.shape[1]
    mod.setup()
    kinds, lo, hi = mod._column_kinds(D)
    scope = mod.scope
    old_min = [None if k == _lib.BOUND_BOTH else scope.get("batch_stats", f"xmin_{i}") for i, k in enumerate(kinds)]
    old_max = [None if k == _lib.BOUND_BOTH else scope.get("batch_stats", f"xmax_{i}") for i, k in enumerate(kinds)]
    pack = lambda vals, fill: torch.tensor(
        [fill if v is None else float(np.asarray(v.cpu() if isinstance(v, torch.Tensor) else v).reshape(-1)[0]) for v in vals],
        dtype=torch.float32, device=dev)
    xmin, xmax = pack(old_min, 0.0), pack(old_max, 0.0)
    sb = _sb_struct(kinds, lo, hi, mod.margin, xmin, xmax)
    _minmax_update(sb, xd, D)
    for i, k in enumerate(kinds):  # bijectors.py:258-260 (store unless initializing)
        if k != _lib.BOUND_BOTH:
            scope.put("batch_stats", f"xmin_{i}", _leaf_out(xmin[i:i + 1].clone(), old_min[i]))
            scope.put("batch_stats", f"xmax_{i}", _leaf_out(xmax[i:i + 1].clone(), old_max[i]))
    ch, keep = _chain_struct(D, 0, [_op_sb(sb)])
    y = torch.empty_like(xd)
    ld = torch.empty(xd.shape[0], dtype=torch.float32, device=dev)
    _run_chain_forward(ch, xd, None, y, ld, acc=False)
    return like_input(y, x), like_input(ld, x)


def coupling_train_forward(mod, x, c):
    """NeuralSplineCoupling.__call__(train=True): batch-moment BatchNorm (bijectors.py:342)."""
    dev = require_cuda()
    xd = to_device_f32(x, dev)
    cd = None if c is None else to_device_f32(c, dev)
    if cd is not None and cd.ndim == 1:
        cd = cd.reshape(-1, 1)
    M, D = xd.shape
    Cdim = 0 if cd is None else cd.shape[1]
    scope = mod.scope
    scale, bias, mean, var, kernels, biases = mod._leaves(scope, D)
    dl = lambda a: to_device_f32(a, dev)
    ra_mean, ra_var = dl(mean).clone(), dl(var).clone()
    F = D - D // 2 + Cdim
    bmean = torch.empty(F, dtype=torch.float32, device=dev)
    bvar = torch.empty(F, dtype=torch.float32, device=dev)
    _bn_batch_stats(xd, cd, D, Cdim, M, bmean, bvar, ra_mean, ra_var)
    bn = scope.child("BatchNorm_0")
    bn.put("batch_stats", "mean", _leaf_out(ra_mean, mean))
    bn.put("batch_stats", "var", _leaf_out(ra_var, var))
    keep = [dl(scale), dl(bias)] + [dl(k) for k in kernels] + [dl(b) for b in biases]
    nl = len(kernels)
    cp = _coupling_struct(mod, keep[0], keep[1], bmean, bvar, keep[2:2 + nl], keep[2 + nl:])
    ch, arr = _chain_struct(D, Cdim, [_op_cp(cp)])
    y = torch.empty_like(xd)
    ld = torch.empty(M, dtype=torch.float32, device=dev)
    _run_chain_forward(ch, xd, cd, y, ld, acc=False)
    return like_input(y, x), like_input(ld, x)


def flow_train_log_prob(flow, x, c):
    """Flow.__call__(train=True), flow.py:45-47 with the bijector in train mode."""
    lib = _lib.load()
    dev = require_cuda()
    with flow.bijector._bound(flow.scope.child("bijector")):
        z, ld = flow.bijector(x, c, True)
    zd, ldd = to_device_f32(z, dev), to_device_f32(ld, dev)
    M, D = zd.shape
    kind, peak = flow.latent._native()
    lp = torch.empty(M, dtype=torch.float32, device=dev)
    gz = torch.empty_like(zd)
    glp = torch.empty(M, dtype=torch.float32, device=dev)
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    _lib.check(lib.zf_flow_loss_grad(stream_ptr(), kind, peak, ptr(zd), ptr(ldd), M, D, float(M), ptr(lp), ptr(gz),
                                     ptr(glp), ptr(acc)), "zf_flow_loss_grad")
    return like_input(lp, x)


# ---------------------------------------------------------------------------------------------
# the train step engine
# ---------------------------------------------------------------------------------------------
class TrainEngine:
    """``step`` of train.py:80-86 on device-resident flat buffers.

    flow.bijector must be a Chain (or a single bijector) made of an optional leading ShiftBounds
    followed by NeuralSplineCoupling / Roll bijectors, which is what the reference builds
    (bijectors.py:374-423).  ``group``: a torch.distributed process group for data-parallel
    training (None: the default group if initialised, else single device).
    """

    def __init__(self, flow, variables, dim: int, cdim: int, *, lr: float = 1e-3, b1: float = 0.9, b2: float = 0.999,
                 eps: float = 1e-8, weight_decay: float = 1e-4, nesterov: bool = True, group=None,
                 micro_batch: int = 1 << 18):
        from .bijectors import Chain, NeuralSplineCoupling, Roll, ShiftBounds

        self.flow = flow
        self.dev = require_cuda()
        self.group = group
        self.hp = dict(lr=lr, b1=b1, b2=b2, eps=eps, weight_decay=weight_decay, nesterov=int(bool(nesterov)))
        self.micro_batch = int(micro_batch)
        self.count = 0
        bij = flow.bijector
        mods = list(bij) if isinstance(bij, Chain) else [bij]
        names = [f"bijectors_{i}" for i in range(len(mods))] if isinstance(bij, Chain) else [None]
        # D comes from the data: the default latent instance is shared between Flow objects and
        # latches its dim only once (flow.py:20, distributions.py:31-32)
        self.D = D = int(dim)
        if D > _lib.ZF_MAX_DIM:
            raise ValueError(f"at most {_lib.ZF_MAX_DIM} columns are supported (got {D})")
        flow.latent._latch_dim(D)
        self.C = int(cdim)
        d = D // 2
        self.F = F = D - d + self.C

        def sub(tree, col, name):
            t = variables.get(col, {}).get("bijector", {})
            return t if name is None else t.get(name, {})

        # ---- groups: a non-Roll bijector and the Rolls that follow it
        self.groups: List[dict] = []
        for mod, name in zip(mods, names):
            if isinstance(mod, Roll):
                if not self.groups:
                    raise NotImplementedError("training a chain that starts with Roll is not supported")
                self.groups[-1]["rolls"].append(int(mod.shift))
            elif isinstance(mod, ShiftBounds):
                if self.groups:
                    raise NotImplementedError("ShiftBounds after another bijector is not supported in training")
                self.groups.append(dict(kind="sb", mod=mod, name=name, rolls=[]))
            elif isinstance(mod, NeuralSplineCoupling):
                self.groups.append(dict(kind="cp", mod=mod, name=name, rolls=[]))
            else:
                raise NotImplementedError(f"cannot train through {type(mod).__name__}")

        # ---- flat parameter buffer with one view per FLAX leaf
        leaves: List[Tuple[dict, str, str, tuple]] = []
        total = 0
        for g in self.groups:
            if g["kind"] != "cp":
                continue
            p = sub(variables, "params", g["name"])
            g["leaf_names"] = [("BatchNorm_0", "scale"), ("BatchNorm_0", "bias")]
            for j in range(len(g["mod"].layers) + 1):
                g["leaf_names"] += [(f"Dense_{j}", "kernel"), (f"Dense_{j}", "bias")]
            g["src"] = p
            for mname, lname in g["leaf_names"]:
                total += int(np.prod(p[mname][lname].shape))
        self.n_params = total
        self.P = torch.empty(total, dtype=torch.float32, device=self.dev)
        self.G = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.mu = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.nu = torch.zeros(total, dtype=torch.float32, device=self.dev)
        off = 0
        for g in self.groups:
            if g["kind"] != "cp":
                continue
            g["pv"], g["gv"] = {}, {}
            for mname, lname in g["leaf_names"]:
                src = g["src"][mname][lname]
                n = int(np.prod(src.shape))
                pv = self.P[off:off + n].view(tuple(src.shape))
                pv.copy_(to_device_f32(src, self.dev))
                g["pv"][(mname, lname)] = pv
                g["gv"][(mname, lname)] = self.G[off:off + n].view(tuple(src.shape))
                off += n
            st = sub(variables, "batch_stats", g["name"])["BatchNorm_0"]
            g["ra_mean"] = to_device_f32(st["mean"], self.dev).clone()
            g["ra_var"] = to_device_f32(st["var"], self.dev).clone()
            g["bmean"] = torch.zeros(F, dtype=torch.float32, device=self.dev)
            g["bvar"] = torch.ones(F, dtype=torch.float32, device=self.dev)
            nl = len(g["mod"].layers) + 1
            ks = [g["pv"][(f"Dense_{j}", "kernel")] for j in range(nl)]
            bs = [g["pv"][(f"Dense_{j}", "bias")] for j in range(nl)]
            # train forward/backward read the BATCH statistics
            g["cp"] = _coupling_struct(g["mod"], g["pv"][("BatchNorm_0", "scale")],
                                       g["pv"][("BatchNorm_0", "bias")], g["bmean"], g["bvar"], ks, bs)
            gr = _lib.ZfCouplingGrads()
            gr.bn_scale = ptr(g["gv"][("BatchNorm_0", "scale")])
            gr.bn_bias = ptr(g["gv"][("BatchNorm_0", "bias")])
            for j in range(nl):
                gr.kernel[j] = ptr(g["gv"][(f"Dense_{j}", "kernel")])
                gr.bias[j] = ptr(g["gv"][(f"Dense_{j}", "bias")])
            g["gr"] = gr
            ops = [_op_cp(g["cp"])] + [_op_roll(s) for s in g["rolls"]]
            g["chain"], g["_keep"] = _chain_struct(D, self.C, ops)
            g["rot"] = sum(g["rolls"])
            g["bn_sums"] = torch.zeros(2 * F, dtype=torch.float64, device=self.dev)
        for g in self.groups:
            if g["kind"] != "sb":
                continue
            mod = g["mod"]
            mod.setup()
            kinds, lo, hi = mod._column_kinds(D)
            st = sub(variables, "batch_stats", g["name"])
            get = lambda key, i, fill: (fill if kinds[i] == _lib.BOUND_BOTH else
                                        float(np.asarray(st[f"{key}_{i}"].cpu() if isinstance(st[f"{key}_{i}"], torch.Tensor)
                                                         else st[f"{key}_{i}"]).reshape(-1)[0]))
            g["xmin"] = torch.tensor([get("xmin", i, 0.0) for i in range(D)], dtype=torch.float32, device=self.dev)
            g["xmax"] = torch.tensor([get("xmax", i, 0.0) for i in range(D)], dtype=torch.float32, device=self.dev)
            g["kinds"] = kinds
            g["sb"] = _sb_struct(kinds, lo, hi, mod.margin, g["xmin"], g["xmax"])
            ops = [_op_sb(g["sb"])] + [_op_roll(s) for s in g["rolls"]]
            g["chain"], g["_keep"] = _chain_struct(D, self.C, ops)
        self._bufs: Dict[int, dict] = {}
        self.lp_sum = torch.zeros(1, dtype=torch.float64, device=self.dev)
        # optimiser step counter in device memory (zf_nadamw_update_dev) + its bias-correction scratch, and the
        # captured single-device steps of small batches (see _step_graphed)
        self._count_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._bias = torch.zeros(4, dtype=torch.float32, device=self.dev)
        self._graphs: Dict[tuple, dict] = {}
        self.use_graphs = os.environ.get("ZF_TRAIN_GRAPHS", "1") != "0"

        # ---- the whole flow as ONE native chain for zf_flow_value_and_grad: couplings point at the flat parameter
        # buffer and at the RUNNING statistics (updated in place by the call)
        ops, self._keep_run = [], []
        cgr = []
        offs = [0]
        for g in self.groups:
            if g["kind"] == "sb":
                ops.append(_op_sb(g["sb"]))
            else:
                nl = len(g["mod"].layers) + 1
                run = _coupling_struct(g["mod"], g["pv"][("BatchNorm_0", "scale")],
                                       g["pv"][("BatchNorm_0", "bias")], g["ra_mean"], g["ra_var"],
                                       [g["pv"][(f"Dense_{j}", "kernel")] for j in range(nl)],
                                       [g["pv"][(f"Dense_{j}", "bias")] for j in range(nl)])
                self._keep_run.append(run)
                ops.append(_op_cp(run))
                cgr.append(g["gr"])
                offs.append(offs[-1] + sum(int(v.numel()) for v in g["gv"].values()))
            ops += [_op_roll(sft) for sft in g["rolls"]]
        self.chain, self._keep_chain = _chain_struct(D, self.C, ops)
        self.n_couplings = len(cgr)
        self._grads_arr = (_lib.ZfCouplingGrads * max(1, len(cgr)))(*cgr)
        self._bucket_off = (C.c_int64 * len(offs))(*offs)
        assert offs[-1] == self.n_params
        self._ws: Dict[int, torch.Tensor] = {}
        dist, world = _dist_world(group)
        self.world = world
        self.use_nccl = world > 1 and dist.get_backend(group) == "nccl"
        self.comm = DpComm(group) if self.use_nccl else None

    # -- buffers sized for a local batch ------------------------------------------------------
    def _buffers(self, M: int) -> dict:
        b = self._bufs.get(M)
        if b is None:
            lib = _lib.load()
            dev, D, F = self.dev, self.D, self.F
            f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
            b = dict(states=[f32(M, D) for _ in self.groups], ld=f32(M), glp=f32(M), ga=f32(M, D), gb=f32(M, D),
                     gh0=f32(M, F), gc=(torch.zeros(M, self.C, dtype=torch.float32, device=dev) if self.C else None))
            mb = min(self.micro_batch, M)
            need = 16
            for g in self.groups:
                if g["kind"] == "cp":
                    need = max(need, int(lib.zf_coupling_backward_workspace_bytes(C.byref(g["cp"]), D, self.C, mb)))
            b["ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
            b["mb"] = mb
            if len(self._bufs) >= 2:  # full and ragged minibatch shapes
                self._bufs.pop(next(iter(self._bufs)))
            self._bufs[M] = b
        return b

    # -- one optimiser step -------------------------------------------------------------------
    def collectives_note(self) -> str:
        if self.world == 1:
            return "none (single device)"
        if self.use_nccl:
            return ("NCCL through the C ABI (zf_dp_*): ShiftBounds min/max as one all-reduce(min), BatchNorm moments fwd+bwd, "
                    "per-coupling gradient buckets on a side stream overlapping the next coupling's backward")
        return "torch.distributed all-reduces between the C-ABI phases (non-NCCL group)"

    def _workspace(self, M: int) -> torch.Tensor:
        ws = self._ws.get(M)
        if ws is None:
            lib = _lib.load()
            need = int(lib.zf_flow_value_and_grad_workspace_bytes(C.byref(self.chain), M, self.micro_batch))
            if need == 0:
                _lib.check(1, "zf_flow_value_and_grad_workspace_bytes")
            if len(self._ws) >= 2:  # full and ragged minibatch shapes
                self._ws.pop(next(iter(self._ws)))
            ws = self._ws[M] = torch.empty(need + 256, dtype=torch.uint8, device=self.dev)
        return ws

    def step(self, x, c=None, *, global_count: Optional[int] = None, update: bool = True, want_gc: bool = False,
             lp_cotangent=None):
        """Loss, gradients and (if ``update``) the optimiser update for the local shard (x, c).

        Returns the device scalar sum of log-probs over the local shard (loss = -sum/global_count);
        with ``want_gc`` also d loss / d c (M, C), the cotangent a conditioner upstream of the flow needs
        (examples/deep_set.ipynb:320-323).  ``lp_cotangent`` (M,) replaces the -1/global_count of the mean loss."""
        lib = _lib.load()
        st = stream_ptr()
        x = to_device_f32(x, self.dev)
        c = None if c is None else to_device_f32(c, self.dev)
        if c is not None and c.ndim == 1:
            c = c.reshape(-1, 1)
        M = x.shape[0]
        if x.shape[1] != self.D or (0 if c is None else c.shape[1]) != self.C:
            raise ValueError(f"expected x (M, {self.D}) and c (M, {self.C})")
        if global_count is None:
            if self.world > 1:
                cnt = torch.tensor([M], dtype=torch.int64, device=self.dev)
                _allreduce(cnt, self.group, "sum")
                global_count = int(cnt.item())
            else:
                global_count = M
        if (self.use_graphs and update and self.world == 1 and lp_cotangent is None and global_count == M
                and 0 < M <= self.GRAPH_MAX_ROWS):
            return self._step_graphed(x, c, M, update, want_gc)
        return self._step_eager(x, c, M, global_count, update, want_gc, lp_cotangent)

    # a step of this many rows or fewer is bound by its ~35 launches per coupling, not by the kernels
    GRAPH_MAX_ROWS = 1 << 16

    def _step_graphed(self, x, c, M: int, update: bool, want_gc: bool):
        """Single-device optimiser step of a small batch as ONE CUDA-graph launch: the first step of a shape runs
        eagerly, the second is captured (inputs staged in fixed buffers; the optimiser's step counter lives on the
        device, so no kernel argument changes between steps) and every later one replays the graph.  With ``want_gc``
        the returned d loss / d c is a copy of the graph's buffer (``value_and_grad`` / ``update=False`` never takes
        this path).  ZF_TRAIN_GRAPHS=0 turns it off."""
        lib = _lib.load()
        key = (M, bool(update), bool(want_gc and self.C))
        e = self._graphs.get(key)
        fresh = e is None
        if fresh:
            if len(self._graphs) >= 4:   # full + ragged minibatch, with and without update
                self._graphs.pop(next(iter(self._graphs)))
            f32 = lambda *sh: torch.empty(*sh, dtype=torch.float32, device=self.dev)
            e = self._graphs[key] = dict(graph=None, launches=0, x=f32(M, self.D), c=f32(M, self.C) if self.C else None,
                                         gc=f32(M, self.C) if key[2] else None, ws=self._workspace(M))
        e["x"].copy_(x)
        if e["c"] is not None:
            e["c"].copy_(c)
        if fresh:
            self._step_eager(e["x"], e["c"], M, M, update, want_gc, None, gc=e["gc"], ws=e["ws"])
        else:
            if e["graph"] is None:
                graph = torch.cuda.CUDAGraph()
                n0 = _lib.launch_count()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    self._step_eager(e["x"], e["c"], M, M, update, want_gc, None, gc=e["gc"], ws=e["ws"], count=False)
                e["launches"] = _lib.launch_count() - n0
                lib.zf_launch_count_add(-e["launches"])   # capturing launched nothing
                e["graph"] = graph
            e["graph"].replay()
            lib.zf_launch_count_add(e["launches"])
            if update:
                self.count += 1
        gc = None if e["gc"] is None else e["gc"].clone()   # the graph's own buffer is overwritten by the next replay
        self._last_gc = gc
        return (self.lp_sum, gc) if want_gc else self.lp_sum

    def _step_eager(self, x, c, M: int, global_count, update: bool, want_gc: bool, lp_cotangent, *, gc=None, ws=None,
                    count: bool = True):
        lib = _lib.load()
        st = stream_ptr()
        self.G.zero_()
        self.lp_sum.zero_()
        if self.world > 1 and not self.use_nccl:
            gc = self._step_phased(x, c, global_count, lp_cotangent)
        else:
            ws = self._workspace(M) if ws is None else ws
            base = (ws.data_ptr() + 255) & ~255
            if gc is None:
                gc = torch.empty(M, self.C, dtype=torch.float32, device=self.dev) if (want_gc and self.C) else None
            comm = self.comm
            aux = comm.aux_stream.cuda_stream if comm is not None else None
            ct = None if lp_cotangent is None else to_device_f32(lp_cotangent, self.dev)
            kind, peak = self.flow.latent._native()
            _lib.check(lib.zf_flow_value_and_grad(
                st, aux, C.byref(self.chain), self._grads_arr, kind, peak, ptr(x), ptr(c), M, float(global_count),
                ptr(ct), None, ptr(self.lp_sum), ptr(gc), None if comm is None else comm.stats,
                None if comm is None else comm.grads, ptr(self.G) if comm is not None else None, self._bucket_off,
                base, ws.numel() - (base - ws.data_ptr()), self.micro_batch), "zf_flow_value_and_grad")
        self._last_gc = gc
        if update:
            h = self.hp
            _lib.check(lib.zf_nadamw_update_dev(st, self.n_params, ptr(self.P), ptr(self.G), ptr(self.mu), ptr(self.nu),
                                                ptr(self._count_dev), ptr(self._bias), h["lr"], h["b1"], h["b2"], h["eps"],
                                                h["weight_decay"], h["nesterov"]), "zf_nadamw_update_dev")
            if count:
                self.count += 1   # host mirror of the device counter
        return (self.lp_sum, gc) if want_gc else self.lp_sum

    def value_and_grad(self, x, c=None, *, global_count: Optional[int] = None):
        """(loss, gradient pytree, d loss / d c) without touching the parameters: what ``jax.value_and_grad`` of
        train.py's ``loss_fn`` returns, plus the cotangent of the conditions."""
        lp_sum, gc = self.step(x, c, global_count=global_count, update=False, want_gc=True)
        n = global_count if global_count is not None else x.shape[0]
        return -float(lp_sum.item()) / n, self.gradients(), gc

    def _step_phased(self, x, c, global_count, lp_cotangent=None):
        """The same step with the phases sequenced from the host and torch.distributed all-reduces in between:
        for process groups that are not NCCL (e.g. gloo in the single-GPU data-parallel parity test)."""
        lib = _lib.load()
        st = stream_ptr()
        M = x.shape[0]
        D, Cd, F = self.D, self.C, self.F
        b = self._buffers(M)
        b["ld"].zero_()
        if b["gc"] is not None:
            b["gc"].zero_()

        # ---- forward, one bijector group at a time (batch statistics couple the samples)
        cur = x
        for gi, g in enumerate(self.groups):
            nxt = b["states"][gi]
            if g["kind"] == "sb":
                _minmax_update(g["sb"], cur, D, self.group)
            else:
                _bn_batch_stats(cur, c, D, Cd, global_count, g["bmean"], g["bvar"], g["ra_mean"], g["ra_var"], self.group)
            g["x_in"] = cur
            _run_chain_forward(g["chain"], cur, c, nxt, b["ld"], acc=True)
            cur = nxt

        # ---- loss and the cotangents of z and of the log-dets
        kind, peak = self.flow.latent._native()
        gy, gx = b["ga"], b["gb"]
        ct = None if lp_cotangent is None else to_device_f32(lp_cotangent, self.dev)
        _lib.check(lib.zf_flow_loss_grad_ct(st, kind, peak, ptr(cur), ptr(b["ld"]), M, D, float(global_count), ptr(ct), None,
                                            ptr(gy), ptr(b["glp"]), ptr(self.lp_sum)), "zf_flow_loss_grad")

        # ---- backward
        for g in reversed(self.groups):
            if g["kind"] != "cp":
                break  # a leading ShiftBounds has no parameters and x needs no cotangent
            _lib.check(lib.zf_coupling_backward(st, C.byref(g["cp"]), C.byref(g["gr"]), D, Cd, ptr(g["x_in"]), ptr(c),
                                                ptr(gy), g["rot"], ptr(b["glp"]), M, ptr(gx), ptr(b["gh0"]),
                                                ptr(g["bn_sums"]), ptr(b["ws"]), b["ws"].numel(), b["mb"]),
                       "zf_coupling_backward")
            _lib.check(lib.zf_bn_param_grads(st, ptr(g["bn_sums"]), F, g["gr"].bn_scale, g["gr"].bn_bias), "zf_bn_param_grads")
            _allreduce(g["bn_sums"], self.group, "sum")
            _lib.check(lib.zf_bn_backward_apply(st, C.byref(g["cp"]), D, Cd, ptr(g["x_in"]), ptr(c), ptr(b["gh0"]),
                                                ptr(g["bn_sums"]), float(global_count), M, ptr(gx), ptr(b["gc"])),
                       "zf_bn_backward_apply")
            gy, gx = gx, gy
        _allreduce(self.G, self.group, "sum")
        return b["gc"]

    # -- FLAX-shaped views of the current state -------------------------------------------------
    def variables(self, as_numpy: bool = False) -> Dict[str, dict]:
        conv = (lambda t: t.detach().cpu().numpy().copy()) if as_numpy else (lambda t: t)
        params: Dict[str, dict] = {}
        stats: Dict[str, dict] = {}
        for g in self.groups:
            if g["kind"] == "cp":
                p: Dict[str, dict] = {}
                for (mname, lname), v in g["pv"].items():
                    p.setdefault(mname, {})[lname] = conv(v)
                s = {"BatchNorm_0": {"mean": conv(g["ra_mean"]), "var": conv(g["ra_var"])}}
            else:
                p = None
                s = {}
                for i, k in enumerate(g["kinds"]):
                    if k != _lib.BOUND_BOTH:
                        s[f"xmin_{i}"] = conv(g["xmin"][i:i + 1])
                        s[f"xmax_{i}"] = conv(g["xmax"][i:i + 1])
            if g["name"] is None:
                if p is not None:
                    params = p
                stats = s
            else:
                if p is not None:
                    params[g["name"]] = p
                stats[g["name"]] = s
        out = {"batch_stats": {"bijector": stats}}
        if params:
            out["params"] = {"bijector": params}
        return out

    def gradients(self) -> Dict[str, dict]:
        """The last step's gradient pytree (views of the flat gradient buffer)."""
        out: Dict[str, dict] = {}
        for g in self.groups:
            if g["kind"] != "cp":
                continue
            p: Dict[str, dict] = {}
            for (mname, lname), v in g["gv"].items():
                p.setdefault(mname, {})[lname] = v
            if g["name"] is None:
                return {"bijector": p}
            out[g["name"]] = p
        return {"bijector": out}

    def snapshot(self):
        """Detached copy of everything ``variables()`` exposes (for best-epoch bookkeeping)."""
        return torch.utils._pytree.tree_map(lambda t: t.clone(), self.variables())
