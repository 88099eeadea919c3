import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, rgbd_b200
from gpu_utils import make_model
B = int(sys.argv[1])
net, sd = make_model(rgbd_b200.ELIC_united, "realistic", 0, precision="bf16")
e = net._program("encoder", B, 512, 640); d = net._program("decoder", B, 8, 10)
print(f"B={B}: encoder plan {e.bytes/2**30:.2f} GB ({e.bytes/B/2**20:.0f} MB/image), decoder plan {d.bytes/2**30:.2f} GB ({d.bytes/B/2**20:.0f} MB/image), torch allocated {torch.cuda.memory_allocated()/2**30:.2f} GB")
