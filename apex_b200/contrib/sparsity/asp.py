"""ASP — automatic 2:4 structured sparsity. Reference: apex/contrib/sparsity/asp.py:27-470: ``init_model_for_pruning`` attaches a mask
buffer to every eligible weight, ``init_optimizer_for_pruning`` patches ``optimizer.step`` so masks are re-applied after every
update, ``compute_sparse_masks`` / ``restore_pruned_weights`` / ``prune_trained_model``. Masks are module buffers, so they travel in
``state_dict`` (reference checkpointing tests). The channel-permutation search (permutation_lib.py + permutation_search_cuda) is
exposed as ``allow_permutation`` but only the identity permutation is implemented in this round."""
from __future__ import annotations

import types

import torch

from .sparse_masklib import create_mask


def eligible_modules(model, whitelist_layer_types, allowed_layer_names, disallowed_layer_names):
    out = []
    for name, mod in model.named_modules():
        if isinstance(mod, whitelist_layer_types) and name not in disallowed_layer_names:
            if allowed_layer_names is not None and name not in allowed_layer_names:
                continue
            out.append((name, mod))
    return out


class ASP:
    __model = None
    __verbosity = 0
    __optimizer = None
    __sparse_parameters = []
    __calculate_mask = None
    __allow_permutation = False
    __permuted = False
    __save_permutation_graph = False
    __permutation_output_dir = "."

    @classmethod
    def init_model_for_pruning(cls, model, mask_calculator="m4n2_1d", verbosity=3,
                               whitelist=(torch.nn.Linear, torch.nn.Conv1d, torch.nn.Conv2d, torch.nn.Conv3d, torch.nn.MultiheadAttention),
                               allowed_layer_names=None, disallowed_layer_names=(), allow_recompute_mask=False, custom_layer_dict=None,
                               allow_permutation=True):
        assert cls.__model is None, "ASP has been initialized already."
        cls.__model, cls.__verbosity, cls.__allow_permutation = model, verbosity, allow_permutation
        cls.__sparse_parameters = []
        cls.__permuted = False
        from .permutation_lib import Permutation

        Permutation.set_permutation_params_from_asp(None, None)
        if isinstance(mask_calculator, str):
            cls.__calculate_mask = lambda p: create_mask(p, mask_calculator).bool()
        else:
            cls.__calculate_mask = mask_calculator
        sparse_names = {torch.nn.Linear: ["weight"], torch.nn.Conv1d: ["weight"], torch.nn.Conv2d: ["weight"], torch.nn.Conv3d: ["weight"],
                        torch.nn.modules.linear.NonDynamicallyQuantizableLinear: ["weight"],
                        torch.nn.MultiheadAttention: ["q_proj_weight", "k_proj_weight", "v_proj_weight", "in_proj_weight"]}
        if custom_layer_dict:
            sparse_names.update(custom_layer_dict)
            whitelist = tuple(whitelist) + tuple(custom_layer_dict.keys())
        for name, mod in eligible_modules(model, tuple(whitelist), allowed_layer_names, tuple(disallowed_layer_names)):
            for p_name in sparse_names.get(type(mod), ["weight"]):
                p = getattr(mod, p_name, None)
                if p is None or not p.requires_grad:
                    continue
                # tensor-core 2:4 needs the pruned (input-channel) dim % 16 == 0 and the other % 8 == 0
                if p.dim() >= 2 and (p.size(0) % 8 != 0 or p.size(1) % 16 != 0):
                    if verbosity >= 3:
                        print(f"[ASP] Auto skipping pruning {name}::{p_name} of size={tuple(p.size())}")
                    continue
                mask = torch.ones_like(p, dtype=torch.bool)
                buf = f"__{p_name}_mma_mask"
                mod.register_buffer(buf, mask)
                pruned = None
                if allow_recompute_mask:
                    pruned = torch.zeros_like(p)
                    mod.register_buffer(f"__{p_name}_mma_pruned_p", pruned)
                cls.__sparse_parameters.append((name, mod, p_name, p, mask, pruned))

    @classmethod
    def set_permutation_saving_params(cls, allow_permutation=True, save_permutation_graph=False, permutation_output_dir="."):
        """Whether compute_sparse_masks() searches channel permutations first, and whether / where the traced graph is dumped
        (reference asp.py:444-475); forwarded to :class:`Permutation`."""
        from .permutation_lib import Permutation

        cls.__allow_permutation = allow_permutation
        cls.__save_permutation_graph = save_permutation_graph
        cls.__permutation_output_dir = permutation_output_dir
        Permutation.set_permutation_saving_params(allow_permutation, save_permutation_graph, permutation_output_dir)

    @classmethod
    def already_init_asp_model(cls):
        return cls.__model is not None

    @classmethod
    def init_optimizer_for_pruning(cls, optimizer):
        assert cls.__optimizer is None, "ASP has initialized optimizer already."
        assert cls.__calculate_mask is not None, "Called ASP.init_optimizer_for_pruning before ASP.init_model_for_pruning."
        cls.__optimizer = optimizer
        optimizer.__step = optimizer.step

        def __step(opt_self, *args, **kwargs):
            with torch.no_grad():
                for _, _, _, p, mask, _ in cls.__sparse_parameters:
                    if p.grad is not None:
                        p.grad.mul_(mask)
            rval = opt_self.__step(*args, **kwargs)
            with torch.no_grad():
                for _, _, _, p, mask, _ in cls.__sparse_parameters:
                    p.mul_(mask)
            return rval

        optimizer.step = types.MethodType(__step, optimizer)

    @classmethod
    def compute_sparse_masks(cls):
        if cls.__allow_permutation and not getattr(cls, "_ASP__permuted", False):
            # search + apply function-preserving channel permutations once, before the first masks (reference asp.py:314-345)
            from .permutation_lib import Permutation

            Permutation.set_permutation_params_from_asp(cls.__model, cls.__sparse_parameters, None, cls.__verbosity)
            Permutation.permute_model(cls.__model, dump_fx_graph=cls.__save_permutation_graph,
                                      save_dumped_fx_graph=(cls.__permutation_output_dir + "/model_offline_permutation_graph.json"
                                                            if cls.__save_permutation_graph else None),
                                      verbosity=cls.__verbosity >= 2)
            cls.__permuted = True
        with torch.no_grad():
            for name, mod, p_name, p, mask, pruned in cls.__sparse_parameters:
                if mask.sum() < mask.numel() and pruned is not None:
                    p.add_(pruned)  # recompute from the dense weights
                mask.set_(cls.__calculate_mask(p).to(mask.dtype)) if mask.shape != p.shape else mask.copy_(cls.__calculate_mask(p))
                if pruned is not None:
                    pruned.copy_(p * (~mask))
                p.mul_(mask)
                if cls.__verbosity >= 2:
                    print(f"[ASP] Enabled {100.0 - 100.0 * mask.float().mean().item():.2f}% sparsity for {name}::{p_name} of size={tuple(p.size())}")

    @classmethod
    def restore_pruned_weights(cls):
        with torch.no_grad():
            for _, _, _, p, mask, pruned in cls.__sparse_parameters:
                if mask.sum() < mask.numel():
                    assert pruned is not None, "Unable to restore dense parameter because allow_recompute_mask == False"
                    p.add_(pruned)
                    mask.fill_(1)
                    pruned.zero_()

    @classmethod
    def is_sparsity_enabled(cls):
        total = sp100 = sp50 = 0
        for _, _, _, _, mask, _ in cls.__sparse_parameters:
            total += 1
            ratio = mask.float().mean().item()
            sp100 += ratio == 1.0
            sp50 += abs(ratio - 0.5) < 1e-6
        if total == sp100:
            return False
        if total == sp50:
            return True
        raise RuntimeError(f"Inconsistent model sparsity: total={total} sp100={sp100} sp50={sp50}")

    @classmethod
    def prune_trained_model(cls, model, optimizer):
        cls.init_model_for_pruning(model, mask_calculator="m4n2_1d", verbosity=2, whitelist=(torch.nn.Linear, torch.nn.Conv2d, torch.nn.MultiheadAttention),
                                   allow_recompute_mask=False)
        cls.init_optimizer_for_pruning(optimizer)
        cls.compute_sparse_masks()

    @classmethod
    def reset(cls):
        """Forget the registered model/optimizer (lets tests run more than once per process)."""
        cls.__model = cls.__optimizer = cls.__calculate_mask = None
        cls.__sparse_parameters = []
        cls.__permuted = False
