#!/bin/bash
python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node 2 distributed_data_parallel.py
