"""N-GPU check of the peer-memory path of row-sharded tables (fused step, CUDA graph):
    torchrun --nproc-per-node 2 tests/dist_p2p_check.py        (wrapped by tests/test_gpu_dist.py)"""
import faulthandler
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "scenario-wise-rec_b200"), HERE, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

faulthandler.dump_traceback_later(150, exit=True)
import dist_parity


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = dist_parity.run(dev, steps=5)
    assert res["ok"] and res["peer_memory_path"] and res["captured_in_graph"], res
    dist.barrier()
    if rank == 0:
        print("DIST_P2P_CHECK_OK", res, flush=True)
    os._exit(0)        # captured step graphs hold NCCL work: destroy_process_group() would wait on it forever


if __name__ == "__main__":
    main()
