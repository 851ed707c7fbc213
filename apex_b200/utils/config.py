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
| ``APEX_B200_LN_FWD_V`` | ``4`` | LayerNorm forward: 16-byte vectors per thread (tuning knob) |
| ``TORCH_SCHED_NUM_STREAMS`` (+ ``_DEBUG``, ``_SKIP_GRAPH_IDS``, ``_REUSE_CUDA_EVENT``, ``_DUMP_CODE``) | ``8`` | torchsched analogue (same names as the reference) |
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
            "TORCH_SCHED_NUM_STREAMS": os.environ.get("TORCH_SCHED_NUM_STREAMS", "8")}
