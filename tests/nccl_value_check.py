"""torchrun worker of tests/test_attack_gpu.py::test_nccl_sharded_gradients_equal_single_process (also run by
bench.py --gpus N through understanding_flow_robustness_b200.attack.nccl_value_check)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from understanding_flow_robustness_b200 import attack
    from understanding_flow_robustness_b200.harness import FlowNetCHarness

    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = FlowNetCHarness(fused_merge=True).to(dev).eval()
    res32 = attack.nccl_value_check(net, dev, rank, world, global_pairs=2 * world, H=128, W=192, p=32)
    net64 = FlowNetCHarness(fused_merge=False).to(dev).eval().double()
    res64 = attack.nccl_value_check(net64, dev, rank, world, global_pairs=2 * world, H=64, W=128, p=16, dtype=torch.float64)
    dist.barrier()
    if rank == 0:
        print(res32)
        print(res64)
        assert res32["ok"] and res64["ok"], (res32, res64)
        print("nccl value check ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
