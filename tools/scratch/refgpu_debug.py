import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "3dgs-deblur_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import ref_bench
v = sys.argv[1]
torch.cuda.set_device(0)
print(v, ref_bench.measure("c2", None, 8, 10, 3, v, breakdown=(v == "cuda")))
