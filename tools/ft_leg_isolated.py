import sys, os, types
sys.path.insert(0, '/root/repo')
os.chdir('/root/repo')
import torch, bench
dev = torch.device('cuda', 0); torch.cuda.set_device(dev)
a = types.SimpleNamespace()
r = bench.finetune_leg(a, dev, 0, 1, None)
print('isolated finetune_leg', r['value'])
r = bench.finetune_leg(a, dev, 0, 1, None, steps=30, warmup=10)
print('isolated finetune_leg, 30 steps after 10 warm-up', r['value'])
