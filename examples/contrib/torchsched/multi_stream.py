"""Multi-stream scheduling of a compiled model with the torchsched backend (reference apex/contrib/torchsched: ``torch.compile(model,
backend="torchsched")``).

    python examples/contrib/torchsched/multi_stream.py                  # forward graph scheduled, interpreted
    python examples/contrib/torchsched/multi_stream.py --codegen        # run the generated multi-stream program instead
    python examples/contrib/torchsched/multi_stream.py --aot --codegen  # forward AND backward graphs scheduled (AOT autograd),
                                                                        # convolution backward split into dgrad / wgrad / bgrad
    TORCH_SCHED_DUMP_CODE=/tmp/ts python ... --codegen                  # also write the programs to /tmp/ts/torchsched/

Works without a GPU (everything then runs on one stream; the plan and the program text are still printed)."""
import argparse
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
from apex_b200.contrib import torchsched  # noqa: E402
from apex_b200.contrib.torchsched import config  # noqa: E402
from apex_b200.contrib.torchsched.inductor import patch_graph_lowering  # noqa: E402


class Block(nn.Module):
    """A residual conv block with two cheap side branches: the main path is the critical path, the branches go to side streams."""

    def __init__(self, c=32):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv2d(c, c, 3, padding=1), nn.Conv2d(c, c, 3, padding=1)
        self.gate, self.skip = nn.Conv2d(c, c, 1), nn.Conv2d(c, c, 1)
        self.norm = nn.LayerNorm(c)

    def forward(self, x):
        h = self.conv2(F.relu(self.conv1(x)))
        g = torch.sigmoid(self.gate(x))
        s = self.skip(x)
        y = (h * g + s).permute(0, 2, 3, 1)
        return self.norm(y).mean((1, 2))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--codegen", action="store_true", help="run generated programs (inductor.patch_graph_lowering)")
    ap.add_argument("--aot", action="store_true", help="schedule backward graphs too (config.aot_autograd)")
    ap.add_argument("--scheme", default="dwb", choices=["dwb", "wbd"], help="order of the convolution-backward pieces")
    ap.add_argument("--streams", type=int, default=config.num_streams)
    args = ap.parse_args()
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    torch.manual_seed(0)
    model = Block().to(dev)
    x = torch.randn(8, 32, 28, 28, device=dev, requires_grad=True)

    backend = torchsched.get_backend("torchsched", scheme=args.scheme)
    patch_graph_lowering(args.codegen)
    with config.patch(aot_autograd=args.aot, num_streams=args.streams):
        compiled = torch.compile(model, backend=backend)
        out = compiled(x)
        out.sum().backward()
    patch_graph_lowering(False)

    ref = model(x)
    print(f"max |compiled - eager| = {(out - ref).abs().max().item():.3e}")
    for sg in backend.graphs:
        print(f"\n=== graph {sg.graph_id} ===")
        print(sg.plan.describe())
        if sg.wrapper_codegen:
            print(sg.program().source)


if __name__ == "__main__":
    main()
