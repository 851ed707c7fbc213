"""SyncBatchNorm fused kernel vs nn.BatchNorm (single GPU: group of one; multi GPU: BatchNorm on the concatenated batch).
Mirrors tests/distributed/synced_batchnorm/{single_gpu_unit_test,two_gpu_unit_test,two_gpu_test_different_batch_size}.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("shape", [(8, 64, 14, 14), (4, 3, 33, 17), (16, 256), (6, 32, 50), (2, 2048, 7, 7), (32, 16, 56, 56)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("channels_last", [False, True])
def test_single_gpu_matches_batchnorm(cuda_dev, shape, dtype, channels_last):
    from apex_b200.parallel import SyncBatchNorm
    if channels_last and len(shape) != 4:
        pytest.skip("channels_last is a 4-D layout")
    torch.manual_seed(0)
    C = shape[1]
    bn = torch.nn.BatchNorm1d(C) if len(shape) < 4 else torch.nn.BatchNorm2d(C)
    bn = bn.to(cuda_dev)
    sbn = SyncBatchNorm(C).to(cuda_dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
        sbn.weight.copy_(bn.weight); sbn.bias.copy_(bn.bias)
    x = (torch.randn(shape, device=cuda_dev) * 2 + 3).to(dtype)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    xr = x.detach().float().requires_grad_(True)
    xs = x.detach().clone().requires_grad_(True)
    dy = torch.randn(shape, device=cuda_dev).to(dtype)
    y_ref = bn(xr)
    y = sbn(xs)
    y_ref.backward(dy.float())
    y.backward(dy)
    ft, bt = (1e-4, 1e-3) if dtype == torch.float32 else (2e-2, 5e-2)
    torch.testing.assert_close(y.float(), y_ref, atol=ft, rtol=ft)
    torch.testing.assert_close(xs.grad.float(), xr.grad, atol=bt, rtol=bt)
    n = x.numel() / C
    torch.testing.assert_close(sbn.weight.grad, bn.weight.grad, atol=bt * n ** 0.5, rtol=bt)
    torch.testing.assert_close(sbn.bias.grad, bn.bias.grad, atol=bt * n ** 0.5, rtol=bt)
    torch.testing.assert_close(sbn.running_mean, bn.running_mean, atol=ft, rtol=ft)
    torch.testing.assert_close(sbn.running_var, bn.running_var, atol=10 * ft, rtol=10 * ft)
    sbn.eval(); bn.eval()
    torch.testing.assert_close(sbn(x).float(), bn(x.float()), atol=ft * 5, rtol=ft * 5)


def test_fuse_relu_and_residual(cuda_dev):
    from apex_b200.parallel import SyncBatchNorm
    torch.manual_seed(0)
    x = torch.randn(8, 32, 10, 10, device=cuda_dev, requires_grad=True)
    z = torch.randn(8, 32, 10, 10, device=cuda_dev, requires_grad=True)
    sbn = SyncBatchNorm(32, fuse_relu=True).to(cuda_dev)
    bn = torch.nn.BatchNorm2d(32).to(cuda_dev)
    xr, zr = x.detach().clone().requires_grad_(True), z.detach().clone().requires_grad_(True)
    y = sbn(x, z)
    yr = torch.relu(bn(xr) + zr)
    dy = torch.randn_like(y)
    y.backward(dy); yr.backward(dy)
    torch.testing.assert_close(y, yr, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(x.grad, xr.grad, atol=1e-3, rtol=1e-3)
    torch.testing.assert_close(z.grad, zr.grad, atol=1e-4, rtol=1e-4)


def test_convert_syncbn_model(cuda_dev):
    from apex_b200.parallel import SyncBatchNorm, convert_syncbn_model
    m = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.ReLU(), torch.nn.Sequential(torch.nn.BatchNorm2d(8)))
    m2 = convert_syncbn_model(m).to(cuda_dev)
    assert isinstance(m2[1], SyncBatchNorm) and isinstance(m2[3][0], SyncBatchNorm)
    m2(torch.randn(2, 3, 8, 8, device=cuda_dev)).sum().backward()


def _two_gpu_case(rank, world, uneven, channels_last):
    import torch.distributed as dist
    from apex_b200.parallel import SyncBatchNorm
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    C = 24
    nb = [4, 7][:world] if uneven else [5] * world
    full = torch.randn(sum(nb), C, 9, 11) * 1.5 + 0.7
    dy_full = torch.randn(sum(nb), C, 9, 11)
    lo = sum(nb[:rank])
    x = full[lo:lo + nb[rank]].to(dev)
    dy = dy_full[lo:lo + nb[rank]].to(dev)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    sbn = SyncBatchNorm(C).to(dev)
    for _ in range(2):  # twice: exercises the double-buffered exchange slots and the running-stat update
        x.grad = None
        y = sbn(x)
        y.backward(dy)
    bn = torch.nn.BatchNorm2d(C).to(dev)
    xr = full.to(dev).requires_grad_(True)
    for _ in range(2):
        xr.grad = None
        yr = bn(xr)
        yr.backward(dy_full.to(dev))
    torch.testing.assert_close(y, yr[lo:lo + nb[rank]], atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(x.grad, xr.grad[lo:lo + nb[rank]], atol=1e-3, rtol=1e-3)
    torch.testing.assert_close(sbn.running_mean, bn.running_mean, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(sbn.running_var, bn.running_var, atol=1e-3, rtol=1e-3)
    gw = sbn.weight.grad.clone()
    dist.all_reduce(gw)
    torch.testing.assert_close(gw, bn.weight.grad, atol=1e-2, rtol=1e-3)


@pytest.mark.parametrize("uneven", [False, True])
@pytest.mark.parametrize("channels_last", [False, True])
def test_two_gpus_match_concatenated_batchnorm(cuda_dev, uneven, channels_last):
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    run_distributed(_two_gpu_case, 2, uneven, channels_last, backend="nccl")


@pytest.mark.parametrize("world", [4, 8])
def test_many_gpus_match_concatenated_batchnorm(cuda_dev, world):
    _need(world)
    from apex_b200.testing.dist_harness import run_distributed
    run_distributed(_two_gpu_case, world, False, True, backend="nccl")


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_functional_ops(cuda_dev, channels_last, dtype):
    """welford_mean_var -> welford_parallel -> batchnorm_forward, reduce_bn -> batchnorm_backward (csrc/syncbn.cpp:71-89)."""
    from apex_b200.parallel import syncbn_ops as S
    torch.manual_seed(1)
    x = (torch.randn(6, 40, 9, 11, device=cuda_dev) * 1.5 + 0.7).to(dtype)
    dy = torch.randn_like(x)
    if channels_last:
        x, dy = x.contiguous(memory_format=torch.channels_last), dy.contiguous(memory_format=torch.channels_last)
    w = torch.rand(40, device=cuda_dev) + 0.5
    b = torch.randn(40, device=cuda_dev)
    mean, var = S.welford_mean_var(x)
    xf = x.float().transpose(0, 1).reshape(40, -1)
    torch.testing.assert_close(mean, xf.mean(1), atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(var, xf.var(1, unbiased=False), atol=1e-4, rtol=1e-3)
    n = x.numel() // 40
    m2, var_u, istd = S.welford_parallel(mean[None], var[None], torch.tensor([n], device=cuda_dev), 1e-5)
    torch.testing.assert_close(var_u, xf.var(1, unbiased=True), atol=1e-4, rtol=1e-3)
    y = S.batchnorm_forward(x, m2, istd, w, b)
    xr = x.float().detach().requires_grad_(True)
    ref = torch.nn.functional.batch_norm(xr, None, None, w, b, True, 0.0, 1e-5)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    torch.testing.assert_close(y.float(), ref, atol=tol, rtol=tol)
    ref.backward(dy.float())
    sdy, sdx, gw, gb = S.reduce_bn(dy, x, m2, istd, w)
    dx = S.batchnorm_backward(dy, x, m2, istd, w, sdy, sdx, torch.tensor([n], device=cuda_dev))
    torch.testing.assert_close(dx.float(), xr.grad, atol=tol * 5, rtol=tol * 5)
    torch.testing.assert_close(gb, dy.float().sum((0, 2, 3)), atol=1e-2, rtol=1e-3)
    z = torch.randn_like(x)
    yz = S.batchnorm_forward_c_last(x, z, m2, istd, w, b, True)
    torch.testing.assert_close(yz.float(), (ref.detach() + z.float()).relu(), atol=tol, rtol=tol)
    g = S.relu_bw_c_last(dy, x, z, m2, istd, w, b)
    torch.testing.assert_close(g.float(), torch.where(ref.detach() + z.float() > 0, dy.float(), torch.zeros_like(dy.float())), atol=0, rtol=0)


def test_ddp_race_condition_two_gpus(cuda_dev):
    """apex_b200.parallel.DistributedDataParallel with a bucket per parameter and 3 all-reduce streams (reference
    tests/distributed/DDP/ddp_race_condition_test.py)."""
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.ddp_race_condition, 2, "cuda", backend="nccl")
