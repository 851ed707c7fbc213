"""Function-preserving channel permutations for 2:4 sparsity. Reference: apex/contrib/sparsity/permutation_lib.py (2,068 LoC:
torch.fx trace -> per-node C/K permutation flags -> sibling / coparent groups -> search -> permute C of the children and K of the
parents, BatchNorm / bias / module attributes in between; behaviour pinned by apex/contrib/sparsity/test/test_permutation_application.py).

Design here: CHANNEL SPACES. The model is symbolically traced and every tensor value in the graph is assigned to a *channel space*
(union-find): the set of values whose channel axis must be permuted together for the network function to stay the same.
  * Linear / Conv / Embedding / MultiheadAttention.out_proj open a new space for their output (they are its PRODUCERS: weight rows,
    bias) and are CONSUMERS of their input's space (weight columns = the dim that gets the 2:4 pattern).
  * elementwise ops (add, mul, residual joins, activations, pooling, dropout) merge / pass through spaces; BatchNorm, InstanceNorm,
    channel LayerNorm, depthwise convolutions and broadcast module attributes are pass-through nodes whose per-channel tensors ride
    along with the space.
  * flatten after a convolution keeps the space but records that every channel now covers `rep` consecutive columns of the next
    Linear (the reference's replicate_sequence case).
  * a concatenation along the channel axis is a list of (space, offset, size) parts: the Linear / Conv / BatchNorm that consumes it is a
    consumer (rider) of EVERY part through the matching slice of its weight, so each input keeps its own permutation (the reference's
    fixup_concats case); parts whose offset or size is not a multiple of 4 would move the 2:4 groups and are frozen instead.
  * a grouped convolution (1 < groups < channels, at least 8 channels per group) cuts its input space into `groups` blocks: the space
    stays permutable, but only by one permutation of the channels-per-group, replicated in every block; ungrouped consumers of the same
    tensor fold their blocks into rows for the search (the reference's init_grouped_conv_permutation_flags). Its output space is frozen.
  * anything that mixes or exposes channel order (graph inputs and outputs, reshapes, matmuls, outputs of grouped convolutions, GroupNorm,
    LocalResponseNorm, slicing, unknown modules and functions) FREEZES the spaces it touches.
Every unfrozen space with at least one prunable consumer is one search problem: the consumers' weights are stacked row-wise (the
reference's sibling group), one permutation is searched (csrc/perm_search.cu), applied to every consumer's input-channel dim and to
every producer / pass-through tensor's channel dim (the reference's coparent group + K_passthru chain). Residual networks and
transformer blocks therefore come out as ONE space for the residual stream and one per hidden FFN / bottleneck width, instead of
being skipped."""
from __future__ import annotations

import operator

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .permutation_search import accelerated_search_for_good_permutation, sum_after_2_to_4

_PRUNABLE = (nn.Linear, nn.Conv1d, nn.Conv2d, nn.Conv3d)
_CONVS = (nn.Conv1d, nn.Conv2d, nn.Conv3d)
_CONV_T = (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d)
_BN_MODULES = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.SyncBatchNorm, nn.InstanceNorm1d, nn.InstanceNorm2d, nn.InstanceNorm3d)
_PASS_MODULES = (nn.ReLU, nn.ReLU6, nn.GELU, nn.SiLU, nn.Sigmoid, nn.Tanh, nn.Dropout, nn.Identity, nn.LeakyReLU, nn.Hardswish, nn.Hardsigmoid,
                 nn.Hardtanh, nn.ELU, nn.SELU, nn.CELU, nn.Mish, nn.Softplus, nn.Softsign)
# modules that act on the trailing (spatial) dims only: channel-preserving for N,C,spatial... values, not for channels-last ones
_SPATIAL_MODULES = (nn.Dropout1d, nn.Dropout2d, nn.Dropout3d, nn.MaxPool1d, nn.MaxPool2d, nn.MaxPool3d, nn.AvgPool1d, nn.AvgPool2d, nn.AvgPool3d,
                    nn.AdaptiveAvgPool1d, nn.AdaptiveAvgPool2d, nn.AdaptiveAvgPool3d, nn.AdaptiveMaxPool1d, nn.AdaptiveMaxPool2d,
                    nn.AdaptiveMaxPool3d, nn.Upsample, nn.ZeroPad2d, nn.ReflectionPad2d, nn.ReplicationPad2d)
_UNARY_FUNCS = {F.relu, F.relu6, F.gelu, F.silu, F.dropout, F.leaky_relu, F.hardswish, F.hardsigmoid, F.hardtanh, F.elu, F.selu, F.mish,
                F.softplus, F.sigmoid, F.tanh, torch.relu, torch.sigmoid, torch.tanh, torch.abs, torch.neg, torch.exp, torch.sqrt, torch.rsqrt,
                torch.square, torch.clamp, torch.clone, operator.neg}
_SPATIAL_FUNCS = {F.max_pool2d, F.avg_pool2d, F.adaptive_avg_pool2d, F.max_pool1d, F.avg_pool1d, F.adaptive_avg_pool1d, F.interpolate}
_BINARY_FUNCS = {operator.add, operator.iadd, operator.sub, operator.isub, operator.mul, operator.imul, operator.truediv, operator.itruediv,
                 torch.add, torch.sub, torch.mul, torch.div, torch.maximum, torch.minimum}
_UNARY_METHODS = {"contiguous", "clone", "detach", "float", "half", "bfloat16", "to", "type", "relu", "sigmoid", "tanh", "clamp", "abs", "neg",
                  "exp", "sqrt", "rsqrt", "square", "pow", "cuda", "cpu", "relu_", "clamp_", "type_as", "requires_grad_"}
_BINARY_METHODS = {"add", "add_", "sub", "sub_", "mul", "mul_", "div", "div_"}
_NON_TENSOR_METHODS = {"size", "dim", "numel", "shape", "ndimension"}
_FLATTEN_FUNCS = {torch.flatten}


def _is_pool_to_1(m):
    if isinstance(m, (nn.AdaptiveAvgPool1d, nn.AdaptiveAvgPool2d, nn.AdaptiveAvgPool3d, nn.AdaptiveMaxPool1d, nn.AdaptiveMaxPool2d, nn.AdaptiveMaxPool3d)):
        o = m.output_size
        return all(v == 1 for v in (o if isinstance(o, (tuple, list)) else (o,)))
    return False


class _Space:
    """One channel space (union-find node). ``consumers``: (module, attr, dim, rep, prunable) -- tensors permuted along their INPUT
    channel dim; ``riders``: (owner, attr, dim) -- producer weights / biases / norm statistics / broadcast attributes permuted along
    their channel dim."""

    # the second row is filled by the search / apply stages of :class:`Permutation` (roots only)
    __slots__ = ("parent", "frozen", "why", "consumers", "riders",
                 "permutation", "skipped", "before", "after", "checked", "search_device", "blocks")

    def __init__(self):
        self.parent, self.frozen, self.why, self.consumers, self.riders = self, False, "", [], []
        self.permutation = self.skipped = self.checked = self.search_device = None
        self.before = self.after = 0.0
        self.blocks = 1

    def find(self):
        s = self
        while s.parent is not s:
            s.parent = s.parent.parent
            s = s.parent
        return s

    def freeze(self, why):
        r = self.find()
        if not r.frozen:
            r.frozen, r.why = True, why

    def union(self, other):
        a, b = self.find(), other.find()
        if a is b:
            return a
        b.parent = a
        a.consumers += b.consumers
        a.riders += b.riders
        if b.frozen and not a.frozen:
            a.frozen, a.why = True, b.why
        b.consumers, b.riders = [], []
        return a


class _Val:
    """A traced tensor value: its channel space, where the channel axis is (1: N,C,spatial...; -1: last dim), the tensor rank when it is
    known (conv outputs), whether the spatial extent is known to be 1, and how many flattened columns each channel covers."""

    __slots__ = ("space", "axis", "rank", "spatial1", "rep", "size")

    def __init__(self, space, axis, rank=None, spatial1=False, rep=1, size=None):
        self.space, self.axis, self.rank, self.spatial1, self.rep, self.size = space, axis, rank, spatial1, rep, size

    def like(self, **kw):
        v = _Val(self.space, self.axis, self.rank, self.spatial1, self.rep, self.size)
        for k, x in kw.items():
            setattr(v, k, x)
        return v


class _Cat:
    """A traced concatenation along the channel axis: ``parts`` = [(value, channels)] in order."""

    __slots__ = ("parts", "axis", "rank")

    def __init__(self, parts, axis, rank):
        self.parts, self.axis, self.rank = parts, axis, rank

    def freeze(self, why):
        for v, _ in self.parts:
            v.space.freeze(why)


class _Slice:
    """Stands in for a module in a consumer / rider entry: the tensors are the [off, off + size) slice along ``dim`` of the owner's."""

    __slots__ = ("owner", "dim", "off", "size")

    def __init__(self, owner, dim, off, size):
        self.owner, self.dim, self.off, self.size = owner, dim, off, size


class _Grouped:
    """Stands in for a grouped convolution (1 < groups < channels) in a consumer entry. Its weight is [K, C / groups, taps...]: group g
    reads the g-th block of C / groups input channels, so the only input permutations it tolerates move channels INSIDE their block, the
    same way in every block — the space is searched over C / groups columns and the result replicated per block (the reference's
    init_grouped_conv_permutation_flags / replicate_sequence, permutation_lib.py:1152-1180,76-85)."""

    __slots__ = ("owner", "groups")

    def __init__(self, owner, groups):
        self.owner, self.groups = owner, groups


def _t(m, attr):
    """The tensor behind a consumer / rider entry (a view for slices: in-place writes reach the parameter)."""
    if isinstance(m, _Slice):
        return getattr(m.owner, attr).narrow(m.dim, m.off, m.size)
    if isinstance(m, _Grouped):
        return getattr(m.owner, attr)
    return getattr(m, attr)


def _base(m, attr):
    return getattr(m.owner if isinstance(m, (_Slice, _Grouped)) else m, attr)


def _tkey(m, attr, dim):
    return (id(_base(m, attr)), dim, m.off if isinstance(m, _Slice) else -1)


def _expand(perm, rep):
    """Channel permutation -> column permutation when each channel covers `rep` consecutive columns (reference: replicate_sequence)."""
    p = torch.as_tensor(perm).long()
    if rep == 1:
        return p
    return (p.view(-1, 1) * rep + torch.arange(rep).view(1, -1)).reshape(-1)


# ---- name helpers of the reference's module-level API (permutation_lib.py:24-85) -------------------------------------------------------
def convert_fx_node_name(fx_node_name: str) -> str:
    """fx spells ``layer1.0.conv1`` as ``layer1_0_conv1``: back to dots (ambiguous for names that contain underscores themselves, which is
    why :func:`node_name_matches` compares without punctuation)."""
    return fx_node_name.replace("_", ".")


def get_node_parent_children(fx_node):
    """(names of the nodes feeding ``fx_node``, names of the nodes reading it), dotted."""
    return ([convert_fx_node_name(n.name) for n in fx_node.all_input_nodes], [convert_fx_node_name(n.name) for n in fx_node.users])


def node_name_matches(node_name: str, module_name: str) -> bool:
    """Does a graph node name denote the module called ``module_name``? Punctuation and case are ignored, and a DDP-wrapped model's
    ``module.`` prefix on the module side is accepted."""
    def squash(name):
        return "".join(ch for ch in name.lower() if ch.isalnum())

    a, b = squash(node_name), squash(module_name)
    return node_name == module_name or a == b or "module." + node_name == module_name or "module" + a == b


def replicate_sequence(sequence, replications: int) -> list:
    """A permutation of C channels applied to ``replications`` consecutive blocks of C (what a flatten in front of a classifier needs)."""
    n = len(sequence)
    return [int(c) + n * r for r in range(replications) for c in sequence]


class Permutation:
    __verbosity = 0
    __seed = 1
    __stats = {"C": 0, "K": 0}
    __sparse_parameters = None
    search_options = {"strategy": "exhaustive", "stripe_group_size": 8, "escape_attempts": 100}

    @classmethod
    def set_identical_seed(cls, identical_seed=1):
        cls.__seed = identical_seed
        torch.manual_seed(identical_seed)

    @classmethod
    def reset_seed(cls):
        """Re-seed before a search so that every rank finds the same permutation (reference permutation_lib.py: reset_seed)."""
        torch.manual_seed(cls.__seed)

    @classmethod
    def set_tcpstore_port(cls, tcpstore_port):
        """The reference synchronises permutations through a TCPStore on this port; here rank 0's result is broadcast over the default
        process group when one exists (sync_permutation), so the port is only recorded."""
        cls.__tcpstore_port = tcpstore_port

    @classmethod
    def set_permutation_saving_params(cls, allow_permutation=False, save_permutation_graph=False, permutation_output_dir="."):
        cls.__allow_permutation = allow_permutation
        cls.__save_permutation_graph = save_permutation_graph
        cls.__permutation_output_dir = permutation_output_dir

    @classmethod
    def set_permutation_params_from_asp(cls, model, sparse_parameters, all_parameters=None, verbosity=0):
        """Hand-over of the ASP state (reference: called from ASP.init_model_for_pruning). When set, only consumers whose weight ASP will
        actually prune drive the search (the reference's one_sparse_sibling case); the others are still permuted along."""
        cls.__model, cls.__verbosity = model, verbosity
        cls.__sparse_parameters = [p for (_, _, _, p, *_rest) in sparse_parameters] if sparse_parameters is not None else None

    @classmethod
    def get_permutation_stats(cls):
        """(tensors permuted along C, tensors permuted along K) by the last permute_model (reference :331)."""
        return cls.__stats["C"], cls.__stats["K"]

    # ------------------------------------------------------------------------------------------------- parameter surgery
    @staticmethod
    def _permute_tensor(t, dim, idx):
        with torch.no_grad():
            t.copy_(t.index_select(dim, idx.to(t.device)))

    @classmethod
    def apply_permutation_in_C_dim(cls, module, perm, attr="weight", dim=1, rep=1):
        """Permute the input channels of a Linear / Conv (dim 1 of the weight; each channel covering `rep` columns after a flatten)."""
        cls._permute_tensor(_t(module, attr), dim, _expand(perm, rep))

    @classmethod
    def apply_permutation_in_K_dim(cls, module, perm):
        """Permute the output channels of a producer: weight rows + bias; BatchNorm: affine parameters and running statistics."""
        idx = torch.as_tensor(perm).long()
        for name in ("weight", "bias", "running_mean", "running_var"):
            t = getattr(module, name, None)
            if t is not None and t.dim() >= 1:
                cls._permute_tensor(t, 0, idx)

    # ------------------------------------------------------------------------------------------------------ graph analysis
    @classmethod
    def build_spaces(cls, model):
        """Trace `model` and return the list of root :class:`_Space` objects (frozen ones included, for reporting)."""
        import torch.fx as fx

        gm = fx.symbolic_trace(model)
        mods = dict(gm.named_modules())
        vals: dict = {}
        spaces: list = []
        first_use: dict = {}   # id(module) -> (input space, output space): a module called twice ties both call sites together

        def poison(m, why):
            """``m`` is called again at a site that cannot be merged with its earlier call(s) (one of them reads a concatenation slice by slice,
            or the module is a grouped convolution): the tensors are shared, so NO call site may be permuted."""
            prev = first_use.get(id(m))
            if prev is None:
                return
            ins, outs = (prev[1], [prev[2]]) if prev[0] == "exclusive" else ([prev[0]], [prev[1]])
            for sp in list(ins) + outs:
                if sp is not None:
                    sp.freeze(why)

        def reuse(m, src_space, out_space):
            prev = first_use.get(id(m))
            if prev is None:
                first_use[id(m)] = (src_space, out_space)
                return False
            if prev[0] == "exclusive":
                why = f"{type(m).__name__} reused at call sites that cannot be merged"
                poison(m, why)
                for sp in (src_space, out_space):
                    if sp is not None:
                        sp.freeze(why)
                return True
            if src_space is not None and prev[0] is not None:
                prev[0].union(src_space)
            if out_space is not None and prev[1] is not None:
                prev[1].union(out_space)
            return True

        def new_space(frozen_why=None):
            s = _Space()
            spaces.append(s)
            if frozen_why:
                s.freeze(frozen_why)
            return s

        def tensor_args(node):
            out = []

            def visit(a):
                if isinstance(a, fx.Node):
                    out.append(a)
                elif isinstance(a, (tuple, list)):
                    for x in a:
                        visit(x)
                elif isinstance(a, dict):
                    for x in a.values():
                        visit(x)

            visit(node.args)
            visit(node.kwargs)
            return out

        def freeze_inputs(node, why):
            for a in tensor_args(node):
                v = vals.get(a)
                if isinstance(v, _Val):
                    v.space.freeze(why)
                elif isinstance(v, _Cat):
                    v.freeze(why)
                elif isinstance(v, tuple):
                    for x in v:
                        if isinstance(x, _Val):
                            x.space.freeze(why)

        def opaque(node, why):
            freeze_inputs(node, why)
            vals[node] = _Val(new_space(why), None)

        def fetch_attr(target):
            obj = gm
            for part in target.split("."):
                obj = getattr(obj, part)
            return obj

        def attr_owner(target):
            parts = target.split(".")
            obj = gm
            for part in parts[:-1]:
                obj = getattr(obj, part)
            real = model
            try:
                for part in parts[:-1]:
                    real = getattr(real, part)
            except AttributeError:
                real = obj
            return real, parts[-1]

        def join(node, operands, why):
            """Elementwise combination of tensor operands (some may be raw get_attr tensors riding on another operand's space)."""
            tvals = [vals.get(a) for a in operands if isinstance(a, fx.Node)]
            if any(isinstance(v, _Cat) for v in tvals):
                return opaque(node, why + " on a concatenation")
            real = [v for v in tvals if isinstance(v, _Val)]
            attrs = [a for a in operands if isinstance(a, fx.Node) and a.op == "get_attr"]
            if not real:
                return opaque(node, why)
            base = real[0]
            for v in real[1:]:
                if v.axis != base.axis or v.rep != base.rep:
                    base.space.freeze("elementwise operands with different layouts")
                    v.space.freeze("elementwise operands with different layouts")
                base.space.union(v.space)
            for a in attrs:
                t = fetch_attr(a.target)
                if not torch.is_tensor(t) or t.numel() == 1:
                    continue
                owner, name = attr_owner(a.target)
                if base.axis == -1:
                    dim = t.dim() - 1
                elif base.axis == 1 and base.rank is not None:
                    dim = t.dim() - (base.rank - 1)
                else:
                    dim = None
                if dim is None or dim < 0:
                    base.space.freeze(f"attribute {a.target} broadcasts ambiguously")
                    continue
                if t.shape[dim] == 1:
                    continue
                base.space.find().riders.append((owner, name, dim, a.target))
            vals[node] = base.like(spatial1=all(v.spatial1 for v in real))

        for node in gm.graph.nodes:
            if node.op == "placeholder":
                vals[node] = _Val(new_space("graph input"), None)
            elif node.op == "get_attr":
                vals[node] = None   # handled where it is used (join); anywhere else the user freezes its other operands
            elif node.op == "output":
                freeze_inputs(node, "graph output")
            elif node.op == "call_module":
                m = mods[node.target]
                src = vals.get(node.args[0]) if node.args and isinstance(node.args[0], fx.Node) else None
                if isinstance(src, _Cat):
                    # consumer / pass-through of a channel concatenation: one sliced entry per part
                    aligned = all(sz % 4 == 0 for _, sz in src.parts)
                    if isinstance(m, _CONVS + (nn.Linear,)) and (not isinstance(m, _CONVS) or m.groups == 1) and aligned and id(m) not in first_use \
                            and src.axis == (1 if isinstance(m, _CONVS) else -1):
                        out = new_space()
                        first_use[id(m)] = ("exclusive", [v.space for v, _ in src.parts], out)
                        off = 0
                        for v, sz in src.parts:
                            v.space.find().consumers.append((_Slice(m, 1, off, sz), "weight", 1, 1, True, node.target, sz))
                            off += sz
                        out.riders.append((m, "weight", 0, node.target))
                        if m.bias is not None:
                            out.riders.append((m, "bias", 0, node.target))
                        is_conv = isinstance(m, _CONVS)
                        vals[node] = _Val(out, 1 if is_conv else -1, m.weight.dim() if is_conv else None,
                                          size=m.out_channels if is_conv else m.out_features)
                    elif isinstance(m, _BN_MODULES) and src.axis == 1 and id(m) not in first_use:
                        first_use[id(m)] = ("exclusive", [v.space for v, _ in src.parts], None)
                        off = 0
                        for v, sz in src.parts:
                            for name in ("weight", "bias", "running_mean", "running_var"):
                                t = getattr(m, name, None)
                                if torch.is_tensor(t) and t.dim() == 1 and not isinstance(t, (nn.parameter.UninitializedParameter, nn.parameter.UninitializedBuffer)):
                                    v.space.find().riders.append((_Slice(m, 0, off, sz), name, 0, node.target))
                            off += sz
                        vals[node] = src
                    elif isinstance(m, _PASS_MODULES) or (isinstance(m, _SPATIAL_MODULES) and src.axis == 1):
                        vals[node] = src
                    else:
                        why = f"{type(m).__name__} at {node.target} on a concatenation"
                        src.freeze(why)
                        poison(m, why)
                        vals[node] = _Val(new_space("consumer of a frozen concatenation"), None)
                    continue
                if isinstance(m, _CONVS + (nn.Linear,)):
                    is_conv = isinstance(m, _CONVS)
                    cin = m.in_channels if is_conv else m.in_features
                    cout = m.out_channels if is_conv else m.out_features
                    groups = m.groups if is_conv else 1
                    rank = (m.weight.dim() if is_conv else None)
                    if not isinstance(src, _Val):
                        opaque(node, "consumer of a non-tensor")
                        continue
                    if groups == 1:
                        want_axis = 1 if is_conv else -1
                        ok = src.axis == want_axis or (not is_conv and src.axis == "flat")
                        if not ok and src.axis is not None:
                            src.space.freeze(f"{node.target}: channel axis mismatch")
                        rep = src.rep if src.axis == "flat" else 1
                        out = new_space()
                        if not reuse(m, src.space, out):
                            src.space.find().consumers.append((m, "weight", 1, rep, True, node.target, cin))
                            out.riders.append((m, "weight", 0, node.target))
                            if m.bias is not None:
                                out.riders.append((m, "bias", 0, node.target))
                        vals[node] = _Val(out, 1 if is_conv else -1, rank, size=cout)
                    elif groups == cin and cout == cin:   # depthwise: per-channel filter, channels pass straight through
                        if not reuse(m, src.space, None):
                            src.space.find().riders.append((m, "weight", 0, node.target))
                            if m.bias is not None:
                                src.space.find().riders.append((m, "bias", 0, node.target))
                        vals[node] = src.like(spatial1=False)
                    elif (src.axis == 1 and cin % groups == 0 and (cin // groups) % 4 == 0 and cin // groups > 4 and id(m) not in first_use):
                        # grouped: the input space stays permutable inside its channel blocks; the output channels are tied to the groups
                        out = new_space(f"output of grouped convolution {node.target}")
                        first_use[id(m)] = ("exclusive", [src.space], out)
                        src.space.find().consumers.append((_Grouped(m, groups), "weight", 1, 1, True, node.target, cin))
                        vals[node] = _Val(out, 1, rank, size=cout)
                    else:
                        poison(m, f"grouped convolution {node.target} called more than once")
                        opaque(node, f"grouped convolution {node.target}")
                elif isinstance(m, _CONV_T):
                    if not isinstance(src, _Val) or m.groups != 1:
                        opaque(node, f"transposed convolution {node.target}")
                        continue
                    out = new_space()
                    if not reuse(m, src.space, out):
                        src.space.find().consumers.append((m, "weight", 0, 1, False, node.target, m.in_channels))
                        out.riders.append((m, "weight", 1, node.target))
                        if m.bias is not None:
                            out.riders.append((m, "bias", 0, node.target))
                    vals[node] = _Val(out, 1, m.weight.dim(), size=m.out_channels)
                elif isinstance(m, nn.Embedding):
                    out = new_space()
                    if not reuse(m, None, out):
                        out.riders.append((m, "weight", 1, node.target))
                    vals[node] = _Val(out, -1, size=m.embedding_dim)
                elif isinstance(m, _BN_MODULES):
                    if not isinstance(src, _Val) or src.axis not in (1, -1):
                        opaque(node, f"normalisation {node.target} on an unknown layout")
                        continue
                    if src.axis == -1 and not isinstance(m, nn.BatchNorm1d):
                        opaque(node, f"normalisation {node.target} on a channels-last value")
                        continue
                    lazy = (nn.parameter.UninitializedParameter, nn.parameter.UninitializedBuffer)
                    for name in ("weight", "bias", "running_mean", "running_var"):
                        t = getattr(m, name, None)
                        if isinstance(t, lazy):
                            src.space.freeze(f"lazy module {node.target} is not materialised")
                        elif torch.is_tensor(t) and t.dim() == 1:
                            src.space.find().riders.append((m, name, 0, node.target))
                    vals[node] = src.like()
                elif isinstance(m, nn.LayerNorm):
                    if not isinstance(src, _Val):
                        opaque(node, "LayerNorm of a non-tensor")
                        continue
                    k = len(m.normalized_shape)
                    if src.axis == -1:    # the channel axis is the last normalised dim: equivariant once gamma / beta are permuted
                        for name in ("weight", "bias"):
                            if getattr(m, name, None) is not None:
                                src.space.find().riders.append((m, name, k - 1, node.target))
                        vals[node] = src.like()
                    elif src.axis == 1 and src.rank is not None:
                        if k == src.rank - 1:   # normalises over [C, spatial...]
                            for name in ("weight", "bias"):
                                if getattr(m, name, None) is not None:
                                    src.space.find().riders.append((m, name, 0, node.target))
                            vals[node] = src.like()
                        elif k < src.rank - 1:  # spatial dims only: channels are untouched
                            vals[node] = src.like()
                        else:
                            opaque(node, f"LayerNorm {node.target} covers the batch dim")
                    else:
                        opaque(node, f"LayerNorm {node.target} on an unknown layout")
                elif isinstance(m, nn.GroupNorm):
                    if isinstance(src, _Val) and src.axis == 1 and m.num_groups in (1, m.num_channels):
                        for name in ("weight", "bias"):
                            if getattr(m, name, None) is not None:
                                src.space.find().riders.append((m, name, 0, node.target))
                        vals[node] = src.like()
                    else:
                        opaque(node, f"GroupNorm {node.target} ties channels into groups")
                elif isinstance(m, nn.Flatten):
                    if isinstance(src, _Val) and src.axis == 1 and m.start_dim == 1 and m.end_dim in (-1, (src.rank or 0) - 1):
                        vals[node] = src.like(axis="flat", rep=1 if src.spatial1 else None)
                    else:
                        opaque(node, f"flatten {node.target}")
                elif isinstance(m, nn.MultiheadAttention):
                    qkv = list(node.args[:3]) + [node.kwargs.get(k) for k in ("query", "key", "value")]
                    ins = [vals.get(a) for a in qkv if isinstance(a, fx.Node)]
                    if len(ins) != 3 or not all(isinstance(v, _Val) and v.axis == -1 for v in ins) or not m._qkv_same_embed_dim:
                        opaque(node, f"MultiheadAttention {node.target} with an unsupported input layout")
                        continue
                    sp = ins[0].space     # q, k and v share in_proj_weight's columns: one permutation for all three inputs
                    for v in ins[1:]:
                        sp.union(v.space)
                    out = new_space()
                    if not reuse(m, sp, out):
                        sp.find().consumers.append((m, "in_proj_weight", 1, 1, True, node.target, m.embed_dim))
                        out.riders.append((m.out_proj, "weight", 0, node.target + ".out_proj"))
                        if m.out_proj.bias is not None:
                            out.riders.append((m.out_proj, "bias", 0, node.target + ".out_proj"))
                    vals[node] = (_Val(out, -1, size=m.embed_dim), None)
                elif isinstance(m, _PASS_MODULES):
                    if isinstance(src, _Val):
                        vals[node] = src.like()
                    else:
                        opaque(node, "pass-through of a non-tensor")
                elif isinstance(m, _SPATIAL_MODULES):
                    if isinstance(src, _Val) and src.axis == 1:
                        vals[node] = src.like(spatial1=src.spatial1 or _is_pool_to_1(m))
                    else:
                        opaque(node, f"spatial module {node.target} on a value whose channel axis is not dim 1")
                else:
                    opaque(node, f"unknown module type {type(m).__name__} at {node.target}")
            elif node.op == "call_function":
                tgt = node.target
                if tgt is operator.getitem:
                    base = vals.get(node.args[0])
                    if isinstance(base, tuple) and isinstance(node.args[1], int) and node.args[1] < len(base):
                        vals[node] = base[node.args[1]]
                    elif isinstance(base, _Val):
                        opaque(node, "tensor slicing")
                    else:
                        vals[node] = None
                elif tgt is getattr:
                    vals[node] = None   # x.shape and friends
                elif tgt in (torch.cat, torch.concat, torch.concatenate):
                    items = node.args[0] if node.args else node.kwargs.get("tensors", ())
                    dim = node.kwargs.get("dim", node.args[1] if len(node.args) > 1 else 0)
                    parts = [vals.get(a) for a in items if isinstance(a, fx.Node)]
                    ok = len(parts) == len(items) and len(parts) > 0 and all(isinstance(v, _Val) and v.size is not None and v.rep == 1 for v in parts)
                    if ok:
                        axis = parts[0].axis
                        ok = all(v.axis == axis for v in parts) and ((axis == 1 and dim == 1) or (axis == -1 and dim == -1))
                    if ok:
                        vals[node] = _Cat([(v, v.size) for v in parts], parts[0].axis, parts[0].rank)
                    else:
                        opaque(node, "concatenation that is not along the channel axis of known-size values")
                elif tgt in _UNARY_FUNCS and node.args and isinstance(vals.get(node.args[0]), _Cat):
                    vals[node] = vals[node.args[0]]
                elif tgt in _UNARY_FUNCS:
                    src = vals.get(node.args[0]) if node.args and isinstance(node.args[0], fx.Node) else None
                    if isinstance(src, _Val):
                        vals[node] = src.like()
                    else:
                        opaque(node, "elementwise function of a non-tensor")
                elif tgt in _SPATIAL_FUNCS:
                    src = vals.get(node.args[0]) if node.args and isinstance(node.args[0], fx.Node) else None
                    if isinstance(src, _Val) and src.axis == 1:
                        osz = node.kwargs.get("output_size", node.args[1] if len(node.args) > 1 else None)
                        to1 = tgt in (F.adaptive_avg_pool2d, F.adaptive_avg_pool1d) and osz in (1, (1, 1), (1,), [1, 1], [1])
                        vals[node] = src.like(spatial1=src.spatial1 or to1)
                    else:
                        opaque(node, f"spatial function {tgt.__name__} on a value whose channel axis is not dim 1")
                elif tgt in _BINARY_FUNCS:
                    join(node, list(node.args[:2]), f"binary {getattr(tgt, '__name__', tgt)}")
                elif tgt in _FLATTEN_FUNCS:
                    src = vals.get(node.args[0])
                    start = node.kwargs.get("start_dim", node.args[1] if len(node.args) > 1 else 0)
                    end = node.kwargs.get("end_dim", node.args[2] if len(node.args) > 2 else -1)
                    if isinstance(src, _Val) and src.axis == 1 and start == 1 and end in (-1, (src.rank or 0) - 1):
                        vals[node] = src.like(axis="flat", rep=1 if src.spatial1 else None)
                    else:
                        opaque(node, "flatten")
                else:
                    opaque(node, f"function {getattr(tgt, '__name__', tgt)}")
            elif node.op == "call_method":
                name = node.target
                src = vals.get(node.args[0]) if node.args and isinstance(node.args[0], fx.Node) else None
                if name in _NON_TENSOR_METHODS:
                    vals[node] = None
                elif name in _UNARY_METHODS and isinstance(src, _Val):
                    vals[node] = src.like()
                elif name in _BINARY_METHODS:
                    join(node, list(node.args[:2]), f"method {name}")
                elif name == "flatten" and isinstance(src, _Val) and src.axis == 1 and (node.args[1:] or (None,))[0] == 1 and len(node.args) <= 2:
                    vals[node] = src.like(axis="flat", rep=1 if src.spatial1 else None)
                elif name in ("view", "reshape") and isinstance(src, _Val) and src.axis == 1 and len(node.args) == 3 and node.args[2] == -1 \
                        and isinstance(node.args[1], fx.Node) and node.args[1].op == "call_method" and node.args[1].target == "size":
                    vals[node] = src.like(axis="flat", rep=1 if src.spatial1 else None)   # x.view(x.size(0), -1)
                else:
                    opaque(node, f"method {name}")

        return [s for s in spaces if s.find() is s]

    @staticmethod
    def _space_size(space):
        for owner, name, dim, _ in space.riders:
            return _t(owner, name).shape[dim]
        return None

    @classmethod
    def _validate(cls, space):
        """-> (ok, C, [resolved consumers]); every tensor's channel dim must agree with the space size."""
        C = cls._space_size(space)
        if C is None or C % 4 != 0:
            return False, C, []
        for owner, name, dim, _ in space.riders:
            t = _t(owner, name)
            if t is None or t.shape[dim] != C:
                return False, C, []
        blocks = {m.groups for m, *_ in space.consumers if isinstance(m, _Grouped)}
        if len(blocks) > 1 or any(C % b for b in blocks):
            return False, C, []         # grouped consumers that cut the channels differently: no common block structure
        space.blocks = blocks.pop() if blocks else 1
        cons = []
        for m, attr, dim, rep, prunable, where, cin in space.consumers:
            w = _t(m, attr)
            if isinstance(m, _Grouped):
                if w.shape[dim] * m.groups != C:
                    return False, C, []
                cons.append((m, attr, dim, 1, prunable, where))
                continue
            if rep is None:   # flattened [C, spatial...]: each channel covers in_features / C consecutive columns
                if w.shape[dim] % C != 0:
                    return False, C, []
                rep = w.shape[dim] // C
            if w.shape[dim] != C * rep:
                return False, C, []
            cons.append((m, attr, dim, rep, prunable, where))
        return True, C, cons

    @classmethod
    def _search_matrix(cls, cons, blocks=1):
        """Stack the prunable consumers' weights into [rows, C / blocks] (spatial taps, flattened repeats and — in a space that a grouped
        convolution cuts into ``blocks`` channel blocks — the blocks of the ungrouped consumers folded into rows)."""
        sparse = cls.__sparse_parameters
        mats = []
        for m, attr, dim, rep, prunable, _ in cons:
            w = _t(m, attr)
            if not prunable or (sparse is not None and not any(_base(m, attr) is p for p in sparse)):
                continue
            w2 = w.detach().movedim(dim, -1).reshape(-1, w.shape[dim]).float()    # [rows, C * rep], channel-major columns
            if rep > 1:
                w2 = w2.reshape(-1, w2.shape[1] // rep, rep).permute(0, 2, 1).reshape(-1, w2.shape[1] // rep)
            if blocks > 1 and not isinstance(m, _Grouped):
                w2 = w2.reshape(-1, w2.shape[1] // blocks)                          # [rows * blocks, C / blocks]
            mats.append(w2)
        return torch.cat(mats, 0) if mats else None

    @staticmethod
    def sync_permutation(perm, device=None):
        """Every rank applies rank 0's permutation (reference: sync_permutations through a TCPStore :1026-1090)."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return perm
        dev = device if dist.get_backend() == "nccl" else "cpu"
        t = torch.as_tensor(perm, dtype=torch.int64, device=dev)
        dist.broadcast(t, 0)
        return [int(i) for i in t.cpu()]

    # ------------------------------------------------------------------------------------------------------------- stages
    # The reference drives the same five stages (permutation_lib.py:196-254: build_fx_graph, init / propagate flags, find_permutations,
    # sync_permutations, apply_permutations); its "flags" are what a _Space carries here (frozen / why / consumers / riders).
    @classmethod
    def build_fx_graph(cls, model, dump_fx_graph=False, save_dumped_fx_graph="./model_fx_graph.json"):
        """(channel spaces of ``model``, success). Tracing failures are not fatal: an untraceable model is left unpermuted, as the
        reference does (:1776-1997). With ``dump_fx_graph`` the space description is written as JSON before anything is searched."""
        try:
            roots = cls.build_spaces(model)
        except Exception as e:  # noqa: BLE001
            if cls.__verbosity:
                print(f"[permutation_lib] model is not fx-traceable ({type(e).__name__}: {e}); skipping channel permutations")
            return [], False
        if dump_fx_graph and save_dumped_fx_graph:
            cls.save_graph_to_json(cls.describe_spaces(roots), save_dumped_fx_graph)
        return roots, True

    @classmethod
    def describe_spaces(cls, roots) -> dict:
        """JSON-able description of the channel spaces: who consumes / rides each, why a space is skipped, and — once the later stages
        ran — the permutation and the kept magnitude before / after."""
        groups = []
        for space in roots:
            d = {"consumers": [c[5] for c in space.consumers], "riders": sorted({f"{r[3]}.{r[1]}" for r in space.riders})}
            if getattr(space, "skipped", None) or space.frozen or not space.consumers:
                d["skipped"] = getattr(space, "skipped", None) or space.why or "no consumers"
            if getattr(space, "permutation", None) is not None:
                d.update(channels=len(space.permutation) * space.blocks, blocks=space.blocks, permutation=space.permutation,
                         kept_magnitude_before=space.before, kept_magnitude_after=space.after)
            groups.append(d)
        return {"groups": groups}

    @classmethod
    def find_permutations(cls, fx_graph) -> int:
        """Search stage: every eligible space gets ``space.permutation`` (None when there is nothing to gain) and the kept-magnitude pair
        ``space.before / space.after``; returns how many spaces found an improving permutation."""
        found = 0
        for space in fx_graph:     # the channel spaces build_fx_graph returned
            space.permutation = space.skipped = None
            if space.frozen or not space.consumers:
                if cls.__verbosity and space.consumers:
                    print(f"[permutation_lib] skipping a space with {len(space.consumers)} consumer(s): {space.why}")
                continue
            ok, C, cons = cls._validate(space)
            if not ok:
                space.skipped = f"dimension mismatch (C = {C})"
                continue
            stacked = cls._search_matrix(cons, space.blocks)
            if stacked is None:
                space.skipped = "no prunable consumer"
                continue
            if stacked.shape[1] <= 4:
                space.skipped = "a single 2:4 group per block: nothing to permute"
                continue
            space.checked, space.search_device = cons, stacked.device
            space.before = float(sum_after_2_to_4(stacked))
            cls.reset_seed()
            perm = [int(i) for i in accelerated_search_for_good_permutation(stacked, cls.search_options, cls.__verbosity)]
            space.after = float(sum_after_2_to_4(stacked[:, torch.as_tensor(perm, device=stacked.device)]))
            space.permutation = perm
            found += space.after > space.before
        return found

    @classmethod
    def sync_permutations(cls, fx_graph) -> None:
        """Distributed stage: rank 0's permutation of every searched space replaces the local one, so that all data-parallel replicas permute
        identically (reference :577-640 broadcasts through a TCPStore / the default group). Ranks search the same weights, so ``before`` is
        identical everywhere and ``after`` is rank 0's."""
        for space in fx_graph:     # the channel spaces build_fx_graph returned
            if getattr(space, "permutation", None) is None:
                continue
            space.permutation = cls.sync_permutation(space.permutation, space.search_device)
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                pair = torch.tensor([space.before, space.after], dtype=torch.float64,
                                    device=space.search_device if dist.get_backend() == "nccl" else "cpu")
                dist.broadcast(pair, 0)
                space.before, space.after = (float(v) for v in pair.cpu())

    @classmethod
    def apply_permutations(cls, fx_graph) -> list:
        """Apply stage: permute the input channels (C) of every consumer and the matching output-side tensors (K: producers' rows, biases,
        norm parameters ... the riders) of every space whose permutation improves the kept magnitude. -> [(n consumers, before, after)]."""
        report = []
        for space in fx_graph:     # the channel spaces build_fx_graph returned
            perm = getattr(space, "permutation", None)
            if perm is None:
                continue
            if space.after <= space.before:
                space.skipped, space.permutation = "no improvement", None
                continue
            full = replicate_sequence(perm, space.blocks)       # the permutation of all C channels (== perm without grouped consumers)
            idx = torch.as_tensor(full, device=space.search_device)
            seen = set()
            for m, attr, dim, rep, _, _ in space.checked:
                if _tkey(m, attr, dim) in seen:
                    continue
                seen.add(_tkey(m, attr, dim))
                cls.apply_permutation_in_C_dim(m, perm if isinstance(m, _Grouped) else full, attr, dim, rep)
                cls.__stats["C"] += 1
            seen = set()
            for owner, name, dim, _ in space.riders:
                if _tkey(owner, name, dim) in seen:      # a module reused at several call sites rides once
                    continue
                seen.add(_tkey(owner, name, dim))
                cls._permute_tensor(_t(owner, name), dim, idx.long())
                cls.__stats["K"] += 1
            report.append((len(space.checked), space.before, space.after))
            if cls.__verbosity:
                print(f"[permutation_lib] space of {len(perm)} channels, {len(space.checked)} consumer(s), {len(seen)} rider tensor(s): kept "
                      f"magnitude {space.before:.3f} -> {space.after:.3f}")
        return report

    @classmethod
    def trace_and_print_raw_fx_graph(cls, model, print_tabular=False, generate_python_code=False):
        """Symbolically trace ``model`` and print its fx graph (optionally as a table / as generated Python); None if it cannot be traced
        (reference :1999-2057)."""
        try:
            traced = torch.fx.symbolic_trace(model)
        except Exception as e:  # noqa: BLE001
            if cls.__verbosity and (not dist.is_available() or not dist.is_initialized() or dist.get_rank() == 0):
                print(f"[print_raw_fx_graph] cannot symbolically trace the model: {type(e).__name__}: {e}")
            return None
        if cls.__verbosity > 1:
            print(traced.graph)
        if print_tabular:
            for n in traced.graph.nodes:
                print(f"{n.op:<14} {n.name:<28} {str(n.target):<40} {[a.name for a in n.all_input_nodes]}")
        if generate_python_code:
            print(traced.code)
        return traced

    @classmethod
    def save_graph_to_json(cls, graph, save_dumped_graph_path_with_name="./model_fx_graph.json"):
        import json

        with open(save_dumped_graph_path_with_name, "w", encoding="utf-8") as f:
            json.dump(graph, f, indent=1)

    # -------------------------------------------------------------------------------------------------------------- driver
    @classmethod
    def permute_model(cls, model, dump_fx_graph=False, save_dumped_fx_graph="./model_permutation_graph.json", verbosity=0):
        """Search and apply a permutation for every unfrozen channel space; returns [(n consumers, magnitude before, after)]."""
        cls.__verbosity = verbosity
        cls.__stats = {"C": 0, "K": 0}
        roots, ok = cls.build_fx_graph(model)
        if not ok:
            return []
        cls.find_permutations(roots)
        cls.sync_permutations(roots)
        report = cls.apply_permutations(roots)
        if dump_fx_graph and save_dumped_fx_graph:   # what was permuted and how, for offline inspection (the reference dumps its annotated fx graph)
            cls.save_graph_to_json(cls.describe_spaces(roots), save_dumped_fx_graph)
        return report
