"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU): a conv -> train-mode BN
block on a batch sharded over the ranks with statistics shared over NCCL must reproduce the single-GPU result on the
whole batch - outputs, running statistics, input gradient and (rank-averaged, as DDP does) parameter gradients.  The
check itself is spiking_diffusion_b200.selfcheck.syncbn_over_nccl, which bench.py also runs under --gpus >= 2."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spiking_diffusion_b200 import selfcheck  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    ok, errs = selfcheck.syncbn_over_nccl()
    if rank == 0:
        print("SYNCBN", "OK" if ok else "FAIL", {k: f"{v:.2e}" for k, v in errs.items()}, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
