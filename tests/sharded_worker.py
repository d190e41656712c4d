"""Worker of tests/test_gpu_pipeline.py::test_config3_two_gpu_shards_equal_one_gpu_bit_for_bit (TEST INFRASTRUCTURE).
Launched by torchrun with 2 ranks: stylises a shared-style batch through parallel.stylize_sharded (frame i -> rank
floor(i*G/B), NCCL all_gather of the finished uint8 frames) and rank 0 saves the gathered batch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALL = ["relu5_1", "relu4_1", "relu3_1", "relu2_1", "relu1_1"]


def make_batch(n, size):
    rng = np.random.default_rng(4321)
    return rng.integers(0, 256, (n, size, size, 3), dtype=np.uint8), rng.integers(0, 256, (1, size, size, 3), dtype=np.uint8)


def main():
    import torch.distributed as dist
    from wct_tf_b200 import parallel
    from wct_tf_b200.wct import WCT
    from wct_tf_b200.weights import make_synthetic_weights
    out_path, n, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wct = WCT(checkpoints=None, relu_targets=ALL, vgg_path=None, device="cuda:%d" % local, weights=make_synthetic_weights(42))
    c, s = make_batch(n, size)
    ct, stt = torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda()
    full = parallel.stylize_sharded(lambda cc, ss: wct.predict_batch(cc, ss, alpha=0.8, return_device=True), ct, stt)
    torch.cuda.synchronize()
    if dist.get_rank() == 0:
        np.save(out_path, full.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
