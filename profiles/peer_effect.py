"""Does enabling CUDA peer access (what NCCL / symmetric memory do at N > 1) slow the single-GPU persistent trainer?
Runs the headline TD bench in ONE process on cuda:0 after (optionally) enabling peer access to cuda:1 with a
device-to-device copy.  usage: python profiles/peer_effect.py [peer]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "peer" in sys.argv[1:]:
    a = torch.ones(1 << 20, device="cuda:1")
    b = a.to("cuda:0")
    torch.cuda.synchronize()
    print("peer access enabled:", torch.cuda.can_device_access_peer(0, 1), float(b.sum()), file=sys.stderr)
import bench  # noqa: E402

sys.argv = ["bench.py", "--steps", "10", "--warmup", "3", "--no-configs", "--no-extras"]
bench.main()
