from .distributed_test_base import DistributedTestBase, GlooDistributedTestBase, NcclDistributedTestBase, UccDistributedTestBase  # noqa: F401
