import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tq100, bench
dev = torch.device("cuda:0")
print(json.dumps(bench.whole_model_leg(torch, tq100, dev, layers=2)))
print(json.dumps(bench.whole_model_leg(torch, tq100, dev, layers=2)))
