"""apex_b200.parallel.DistributedDataParallel / Reducer on two gloo ranks (CPU): bucketing, delayed all-reduce, pre-division,
and the reference's race-by-construction check (tests/distributed/DDP/ddp_race_condition_test.py)."""
import pytest

from apex_b200.testing.dist_harness import run_distributed
from tests import _dist_cases as cases


@pytest.mark.parametrize("delay,message_size,predivide", [(False, 1, 1.0), (False, 50, 2.0), (True, 10000000, 1.0)])
def test_ddp_two_ranks_gloo(delay, message_size, predivide):
    run_distributed(cases.ddp_matches_manual_allreduce, 2, "cpu", delay, message_size, predivide, backend="gloo")


def test_ddp_race_condition_gloo():
    run_distributed(cases.ddp_race_condition, 2, "cpu", backend="gloo")


def test_spatial_bottleneck_gloo():
    run_distributed(cases.spatial_bottleneck_matches_full, 2, "cpu", backend="gloo")


def test_spatial_bottleneck_function_gloo():
    run_distributed(cases.spatial_bottleneck_function_matches_full, 2, "cpu", backend="gloo")


@pytest.mark.parametrize("uneven,fuse_relu", [(False, False), (True, True)])
def test_syncbn_generic_path_gloo(uneven, fuse_relu):
    run_distributed(cases.syncbn_generic_matches_concatenated_batchnorm, 2, "cpu", uneven, fuse_relu, backend="gloo")


def test_halo_exchangers_three_ranks_gloo():
    run_distributed(cases.halo_exchangers_match_slices_of_the_full_tensor, 3, "cpu", backend="gloo")


def test_group_batchnorm_four_ranks_gloo():
    run_distributed(cases.group_batchnorm_spans_only_its_group, 4, "cpu", backend="gloo")


@pytest.mark.parametrize("delay", [False, True])
def test_ddp_unused_parameters_gloo(delay):
    run_distributed(cases.ddp_reduces_when_some_parameters_get_no_gradient, 2, "cpu", delay, backend="gloo")


def test_reducer_gloo():
    run_distributed(cases.reducer_averages_gradients_and_broadcasts_parameters, 3, "cpu", backend="gloo")


def test_ddp_option_matrix_gloo():
    run_distributed(cases.ddp_option_matrix, 2, "cpu", backend="gloo")


def test_cudnn_gbn_lib_raw_entry_points_two_ranks_gloo():
    run_distributed(cases.cudnn_gbn_lib_group_of_two, 2, "cpu", backend="gloo")


def test_permutation_sync_gloo():
    run_distributed(cases.permutation_sync_uses_rank0, 2, "cpu", backend="gloo")
