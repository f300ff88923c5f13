"""Numerics self-test of the tensor-parallel engine (run with torch.distributed.run on >= 2 GPUs): prints one JSON line on rank 0 and
exits non-zero on a mismatch. The test itself is petals_b200/parallel/selftests.py:tp_selftest (bench.py --gpus N runs it too)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from petals_b200.parallel.selftests import tp_selftest

    local = int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", device_id=dev)
    report = tp_selftest(dev, os.environ.get("TP_SELFTEST_MODEL", "llama-tiny"))
    if report:
        print(json.dumps(report))
    dist.destroy_process_group()
    if report and report["tp_selftest"] != "ok":
        sys.exit(1)


if __name__ == "__main__":
    main()
