"""unittest base classes for multi-process tests on one node (reference apex/distributed_testing/distributed_test_base.py:25-131:
``DistributedTestBase`` over torch's ``MultiProcessTestCase``, NCCL and UCC flavours). Here the process management is
:func:`apex_b200.testing.dist_harness.run_distributed` (spawn + file:// rendezvous, traceback of the first failing rank); a test
method decorated with :func:`distributed` runs once per rank with ``self.rank`` / ``self.world_size`` set."""
from __future__ import annotations

import functools
import unittest

import torch

from ..testing.dist_harness import run_distributed


def _run_method(rank, world, cls, name):
    case = cls(name)
    case.rank, case.world_size = rank, world
    case.setUp()
    try:
        getattr(case, name).__wrapped__(case)
    finally:
        case.tearDown()


def distributed(fn):
    """Mark a TestCase method as multi-process: the parent spawns ``world_size`` ranks that each run the undecorated body."""

    @functools.wraps(fn)
    def wrapper(self):
        run_distributed(_run_method, self.world_size, type(self), fn.__name__, backend=self.DISTRIBUTED_BACKEND)

    wrapper.__wrapped__ = fn
    return wrapper


class DistributedTestBase(unittest.TestCase):
    DISTRIBUTED_BACKEND = None  # None = nccl when every rank can own a GPU, else gloo
    rank = 0

    @property
    def world_size(self) -> int:
        return getattr(self, "_world_size", None) or min(max(torch.cuda.device_count(), 2), 4)

    @world_size.setter
    def world_size(self, v: int) -> None:
        self._world_size = v

    @property
    def init_method(self) -> str:
        """Rendezvous of the running process group (the harness uses a file:// store private to each test, 127.0.0.1 otherwise)."""
        import os

        return os.environ.get("APEX_B200_DIST_INIT_METHOD", "env://")

    @property
    def destroy_pg_upon_exit(self) -> bool:
        """The reference overrides torch's default so that the group outlives the test body (:46-48); the harness tears it down itself."""
        return False


class NcclDistributedTestBase(DistributedTestBase):
    DISTRIBUTED_BACKEND = "nccl"


class GlooDistributedTestBase(DistributedTestBase):
    DISTRIBUTED_BACKEND = "gloo"


@unittest.skipUnless(torch.distributed.is_available() and getattr(torch.distributed, "is_ucc_available", lambda: False)(),
                     "this torch build has no UCC backend")
class UccDistributedTestBase(DistributedTestBase):
    DISTRIBUTED_BACKEND = "ucc"
