from .distributed_test_base import DistributedTestBase, NcclDistributedTestBase, GlooDistributedTestBase  # noqa: F401
