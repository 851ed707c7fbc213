"""``nccl_p2p`` halo primitives (reference apex/contrib/csrc/nccl_p2p/nccl_p2p_cuda.cu:34-205: a second NCCL communicator with
grouped ncclSend/ncclRecv). torch.distributed's batched P2P ops issue the same grouped send/recv on the existing communicator."""
from .nccl_p2p import add_delay, get_unique_nccl_id, init_nccl_comm, left_right_halo_exchange, left_right_halo_exchange_inplace  # noqa: F401
