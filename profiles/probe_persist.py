import os, sys, time
os.environ["FFB_PD_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(sys.path[0], "tests"))
import numpy as np, torch
from util import load_case
from faceformer_b200.engine import Engine
g = load_case("seq2seq_single64")
e = Engine(g["cfg"], g["mode"], 0); e.load_state_dict(g["sd"])
b = g["batch"]
coords = torch.from_numpy(b["input"]).cuda().flatten(2); mask = torch.from_numpy(b["input_mask"]).cuda(); ni = torch.from_numpy(b["num_input"]).cuda()
for i in range(2):
    torch.cuda.synchronize(); t = time.time()
    pred, steps = e.forward_eval(coords, mask, ni)
    torch.cuda.synchronize(); print("ms", (time.time() - t) * 1e3, "steps", steps, file=sys.stderr)
