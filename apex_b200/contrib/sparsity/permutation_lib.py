"""Function-preserving channel permutations for 2:4 sparsity. Reference: apex/contrib/sparsity/permutation_lib.py (2,068 LoC:
torch.fx trace -> sibling / parent / child groups -> search -> permute C of consumers and K of producers, plus BN / bias).

Same idea, compact: the model is symbolically traced; for every prunable consumer (Linear / Conv whose INPUT-channel dim gets the
2:4 pattern) we walk back through channel-preserving nodes (activations, dropout, BatchNorm / LayerNorm-free elementwise ops) to
the layer(s) that PRODUCE those channels. Consumers that share a producer are siblings: their weights are stacked row-wise and one
permutation is searched for the group (csrc/perm_search.cu). The permutation is applied to the consumers' input channels, and the
inverse bookkeeping (output channels of the producer: weight rows, bias, BatchNorm affine + running stats in between) keeps the
network function unchanged. Anything the walk cannot prove safe (residual adds, reshapes, graph inputs) is left alone."""
from __future__ import annotations

import operator

import torch
import torch.nn as nn
import torch.nn.functional as F

from .permutation_search import accelerated_search_for_good_permutation, sum_after_2_to_4

_PASS_MODULES = (nn.ReLU, nn.ReLU6, nn.GELU, nn.SiLU, nn.Sigmoid, nn.Tanh, nn.Dropout, nn.Dropout2d, nn.Identity, nn.LeakyReLU, nn.Hardswish,
                 nn.MaxPool2d, nn.AvgPool2d, nn.AdaptiveAvgPool2d)
_BN_MODULES = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)
_PASS_FUNCS = (F.relu, F.gelu, F.silu, F.dropout, torch.relu, torch.sigmoid, torch.tanh, F.leaky_relu, F.hardswish)
_PRUNABLE = (nn.Linear, nn.Conv1d, nn.Conv2d)


class Permutation:
    __verbosity = 0
    __seed = 1
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
        """The reference synchronises permutations through a TCPStore on this port; here every rank runs the same seeded search, so the
        port is only recorded."""
        cls.__tcpstore_port = tcpstore_port

    @classmethod
    def set_permutation_saving_params(cls, allow_permutation=True, save_permutation_graph=False, permutation_output_dir="."):
        cls.__allow_permutation = allow_permutation
        cls.__save_permutation_graph = save_permutation_graph
        cls.__permutation_output_dir = permutation_output_dir

    @classmethod
    def set_permutation_params_from_asp(cls, model, sparse_parameters, all_parameters=None, verbosity=0):
        """Hand-over of the ASP state (reference: called from ASP.init_model_for_pruning); the graph search works from the model alone."""
        cls.__model, cls.__sparse_parameters, cls.__verbosity = model, sparse_parameters, verbosity

    # ------------------------------------------------------------------------------------------------- parameter surgery
    @staticmethod
    def apply_permutation_in_C_dim(module, perm):
        """Permute the input channels (dim 1 of the weight) of a Linear / Conv."""
        idx = torch.as_tensor(perm, device=module.weight.device).long()
        with torch.no_grad():
            module.weight.copy_(module.weight.index_select(1, idx))

    @staticmethod
    def apply_permutation_in_K_dim(module, perm):
        """Permute the output channels of a producer: weight rows + bias; BatchNorm: affine parameters and running statistics."""
        with torch.no_grad():
            for name in ("weight", "bias", "running_mean", "running_var"):
                t = getattr(module, name, None)
                if t is not None and t.dim() >= 1:
                    t.copy_(t.index_select(0, torch.as_tensor(perm, device=t.device).long()))

    # ------------------------------------------------------------------------------------------------------ graph analysis
    @classmethod
    def build_groups(cls, model):
        """-> list of (consumers [modules], producers [modules], in-between BatchNorms [modules])."""
        import torch.fx as fx

        gm = fx.symbolic_trace(model)
        mods = dict(gm.named_modules())

        def producers_of(node, bns, seen):
            """Walk up from `node`; returns the list of producer modules, or None if the path is not provably channel-preserving."""
            if node in seen:
                return []
            seen.add(node)
            if node.op == "call_module":
                m = mods[node.target]
                if isinstance(m, _PRUNABLE):
                    if isinstance(m, (nn.Conv1d, nn.Conv2d)) and m.groups != 1:
                        return None
                    return [(node, m)]
                if isinstance(m, _BN_MODULES):
                    bns.append(m)
                    return producers_of(node.args[0], bns, seen)
                if isinstance(m, _PASS_MODULES):
                    return producers_of(node.args[0], bns, seen)
                return None
            if node.op == "call_function" and node.target in _PASS_FUNCS:
                return producers_of(node.args[0], bns, seen)
            if node.op == "call_function" and node.target in (operator.add, torch.add):
                out = []
                for a in node.args[:2]:
                    if not isinstance(a, fx.Node):
                        continue
                    r = producers_of(a, bns, seen)
                    if r is None:
                        return None
                    out += r
                return out
            return None

        # consumer -> (producer nodes, bns); then merge consumers that share any producer (siblings)
        info = []
        for node in gm.graph.nodes:
            if node.op == "call_module" and isinstance(mods[node.target], _PRUNABLE):
                m = mods[node.target]
                if isinstance(m, (nn.Conv1d, nn.Conv2d)) and m.groups != 1:
                    continue
                bns: list = []
                prods = producers_of(node.args[0], bns, set())
                if prods:
                    info.append((m, prods, bns))
        # every user path of a producer must end in consumers of the same group, otherwise permuting its outputs changes the function
        consumer_inputs = {}
        for m, prods, bns in info:
            for pn, pm in prods:
                consumer_inputs.setdefault(pn, []).append(m)

        def escapes(pn):
            """True if the producer's output reaches anything other than pass-through nodes and prunable consumers."""
            stack, seen = list(pn.users), set()
            while stack:
                u = stack.pop()
                if u in seen:
                    continue
                seen.add(u)
                if u.op == "call_module":
                    mm = mods[u.target]
                    if isinstance(mm, _PRUNABLE):
                        if isinstance(mm, (nn.Conv1d, nn.Conv2d)) and mm.groups != 1:
                            return True
                        continue
                    if isinstance(mm, _PASS_MODULES + _BN_MODULES):
                        stack += list(u.users)
                        continue
                    return True
                if u.op == "call_function" and (u.target in _PASS_FUNCS or u.target in (operator.add, torch.add)):
                    stack += list(u.users)
                    continue
                return True
            return False

        # union-find over consumers sharing producers
        parent = list(range(len(info)))

        def find(i):
            while parent[i] != i:
                parent[i] = parent[parent[i]]
                i = parent[i]
            return i

        owner = {}
        for i, (m, prods, bns) in enumerate(info):
            for pn, pm in prods:
                if pn in owner:
                    parent[find(i)] = find(owner[pn])
                else:
                    owner[pn] = i
        groups = {}
        for i, (m, prods, bns) in enumerate(info):
            g = groups.setdefault(find(i), ([], {}, []))
            g[0].append(m)
            for pn, pm in prods:
                g[1][pn] = pm
            for b in bns:
                if all(b is not x for x in g[2]):
                    g[2].append(b)
        out = []
        for cons, prods, bns in groups.values():
            if any(escapes(pn) for pn in prods):
                continue
            C = cons[0].weight.shape[1]
            if any(c.weight.shape[1] != C for c in cons) or any(p.weight.shape[0] != C for p in prods.values()) or C % 4 != 0:
                continue
            out.append((cons, list(prods.values()), bns))
        return out

    # -------------------------------------------------------------------------------------------------------------- driver
    @classmethod
    def permute_model(cls, model, dump_fx_graph=False, save_dumped_fx_graph=None, verbosity=0):
        """Search and apply a permutation for every safe sibling group; returns [(consumer names..., magnitude before, after)]."""
        cls.__verbosity = verbosity
        try:
            groups = cls.build_groups(model)
        except Exception as e:  # untraceable model: leave it unpermuted, like the reference does on trace failure
            if verbosity:
                print(f"[permutation_lib] model is not fx-traceable ({type(e).__name__}: {e}); skipping channel permutations")
            return []
        report = []
        names = {m: n for n, m in model.named_modules()}
        dumped = []
        for cons, prods, bns in groups:
            mats = [c.weight.detach().reshape(c.weight.shape[0], c.weight.shape[1], -1).permute(0, 2, 1).reshape(-1, c.weight.shape[1]) for c in cons]
            stacked = torch.cat(mats, 0).float()
            before = float(sum_after_2_to_4(stacked))
            perm = accelerated_search_for_good_permutation(stacked, cls.search_options, verbosity)
            after = float(sum_after_2_to_4(stacked[:, torch.as_tensor(perm, device=stacked.device)]))
            if after <= before:
                continue
            for c in cons:
                cls.apply_permutation_in_C_dim(c, perm)
            for p in prods + bns:
                cls.apply_permutation_in_K_dim(p, perm)
            report.append((len(cons), before, after))
            dumped.append({"consumers": [names.get(c, "?") for c in cons], "producers": [names.get(m, "?") for m in prods],
                           "norms": [names.get(m, "?") for m in bns], "permutation": [int(i) for i in perm], "kept_magnitude_before": before,
                           "kept_magnitude_after": after})
            if verbosity:
                print(f"[permutation_lib] group of {len(cons)} consumer(s): kept magnitude {before:.3f} -> {after:.3f}")
        if dump_fx_graph and save_dumped_fx_graph:   # what was permuted and how, for offline inspection (the reference dumps its annotated fx graph)
            import json

            with open(save_dumped_fx_graph, "w") as f:
                json.dump({"groups": dumped}, f, indent=1)
        return report
