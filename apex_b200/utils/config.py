"""Run-time knobs in one place. Like the reference (constructor kwargs plus a handful of environment variables — SURVEY.md §5.6:
``APEX_GROUP_NORM_*_SM_MARGIN``, ``TORCH_SCHED_*``, ``NCCL_NVLS_ENABLE``) there is no global config object; this module only
documents and parses the environment variables the library reads.

| variable | default | meaning |
|---|---|---|
| ``APEX_B200_DIST_NVLS`` | ``auto`` | fused ZeRO step: ``1`` force NVSwitch multimem (NVLS), ``0`` force P2P pull/push, ``auto`` = NVLS from 4 ranks |
| ``APEX_B200_DIST_HYBRID`` | ``1.0`` | fused ZeRO step at >= 4 ranks: fraction of the buckets that goes through the NVSwitch (multimem); the rest runs as a co-resident P2P kernel |
| ``APEX_B200_DIST_OVERLAP_CTAS`` | ``64`` | CTAs of a per-bucket reduce-scatter launched by the overlap_grad_sync hooks while backward is running |
| ``APEX_B200_SYNCBN_SM_MARGIN`` | ``32`` | SMs the multi-GPU SyncBN kernel leaves free for concurrent kernels (NCCL) of the same process |
| ``APEX_B200_GEMM_1CTA`` | unset | force the single-CTA tcgen05 GEMM (the 2-CTA ``cta_group::2`` kernel is the default for M, N > 128) |
| ``APEX_B200_GEMM_NO3D`` | unset | stage MN-major GEMM operands with 2 / 4 two-dimensional TMA boxes per stage instead of one 3-D box (A/B knob) |
| ``APEX_B200_TF32`` | unset | ``1`` / ``0``: fp32 GEMM operands may / may not be multiplied as TF32 on the tcgen05 kernel (default: ``torch.backends.cuda.matmul.allow_tf32``) |
| ``APEX_B200_LN_FWD_V`` | ``4`` | LayerNorm forward: 16-byte vectors per thread (tuning knob) |
| ``APEX_B200_LN_GB_MODE`` | ``2`` | LayerNorm forward gamma / beta staging: ``2`` raw 16-bit words in shared memory, ``1`` fp32 copies, ``0`` global loads |
| ``APEX_B200_LN_BWD_CLUSTER`` | ``1`` | LayerNorm backward: split rows wider than 2 x 512 vectors over a 2-CTA cluster when the halves fill their threads (``2``: whenever possible, ``0``: never) |
| ``APEX_B200_GN_STREAM_MIN_MB`` | unset | GroupNorm: activation size (MB) from which the two-pass streaming kernels replace the slab kernels (default rule: backward slabs >= 300 KB) |
| ``APEX_B200_MT_CHUNK`` / ``APEX_B200_MT_GRID_MULT`` | ``65536`` / ``12`` | multi-tensor engine: elements per work item / CTAs per SM of the persistent grid |
| ``TORCH_SCHED_NUM_STREAMS`` (+ ``_DEBUG``, ``_SKIP_GRAPH_IDS``, ``_REUSE_CUDA_EVENT``, ``_DUMP_CODE``) | ``8`` | torchsched analogue (same names as the reference) |
| ``TORCH_SCHED_CODEGEN`` / ``TORCH_SCHED_AOT`` | ``0`` / ``0`` | torchsched: run the generated multi-stream program instead of the interpreter / schedule forward AND backward graphs through AOT autograd |
"""
from __future__ import annotations

import os


def dist_nvls_policy() -> str:
    v = os.environ.get("APEX_B200_DIST_NVLS", "auto").lower()
    return {"1": "on", "true": "on", "on": "on", "0": "off", "false": "off", "off": "off"}.get(v, "auto")


def gemm_force_1cta() -> bool:
    return os.environ.get("APEX_B200_GEMM_1CTA") is not None


def syncbn_sm_margin() -> int:
    try:
        return max(0, int(os.environ.get("APEX_B200_SYNCBN_SM_MARGIN", "32")))
    except ValueError:
        return 32


def flags() -> dict:
    """Every knob with its current value (for logging at start-up)."""
    return {"APEX_B200_DIST_NVLS": dist_nvls_policy(), "APEX_B200_GEMM_1CTA": gemm_force_1cta(),
            "APEX_B200_LN_FWD_V": os.environ.get("APEX_B200_LN_FWD_V", "4"), "APEX_B200_SYNCBN_SM_MARGIN": syncbn_sm_margin(),
            "APEX_B200_DIST_HYBRID": os.environ.get("APEX_B200_DIST_HYBRID", "1.0"),
            "APEX_B200_DIST_OVERLAP_CTAS": os.environ.get("APEX_B200_DIST_OVERLAP_CTAS", "64"),
            "APEX_B200_TF32": os.environ.get("APEX_B200_TF32"), "APEX_B200_GEMM_NO3D": os.environ.get("APEX_B200_GEMM_NO3D") is not None,
            "APEX_B200_LN_GB_MODE": os.environ.get("APEX_B200_LN_GB_MODE", "2"), "APEX_B200_LN_BWD_CLUSTER": os.environ.get("APEX_B200_LN_BWD_CLUSTER", "1"),
            "APEX_B200_GN_STREAM_MIN_MB": os.environ.get("APEX_B200_GN_STREAM_MIN_MB"),
            "APEX_B200_MT_CHUNK": os.environ.get("APEX_B200_MT_CHUNK", "65536"), "APEX_B200_MT_GRID_MULT": os.environ.get("APEX_B200_MT_GRID_MULT", "12"),
            "TORCH_SCHED_NUM_STREAMS": os.environ.get("TORCH_SCHED_NUM_STREAMS", "8")}
