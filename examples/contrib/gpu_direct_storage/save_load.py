"""Write a tensor to a file and read it back with GDSFile (reference apex/contrib/examples/gpu_direct_storage/example_{save,load}.py).
    python examples/contrib/gpu_direct_storage/save_load.py [path]"""
import sys

import torch

from apex_b200.contrib.gpu_direct_storage import GDSFile

path = sys.argv[1] if len(sys.argv) > 1 else "/tmp/apex_b200_gds_example.bin"
device = "cuda" if torch.cuda.is_available() else "cpu"
x = torch.randn(1024, 1024, device=device)
with GDSFile(path, "w") as f:
    f.save_data(x)
y = torch.empty_like(x)
with GDSFile(path, "r") as f:
    f.load_data(y)
print("round trip ok:", torch.equal(x, y), "bytes:", x.numel() * x.element_size(), "device:", device)
