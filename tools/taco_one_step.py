"""One eager Tacotron train step at Cfg-3 shapes (for `ncu` launch lists): python tools/taco_one_step.py [T_out]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hparams import hparams
from t2_import import t2
from tools.bench_taco import batch

hp = hparams.copy()
hp.parse("predict_linear=False")
B, T_in = 32, 160
T_out = int(sys.argv[1]) if len(sys.argv) > 1 else 800
inputs, lens, mel, stop = batch(hp, B, T_in, T_out)
model = t2.tacotron.Tacotron(hp, B, T_in, T_out)
model.init_variables(seed=5339)
args = (inputs.int().cuda(), lens.int().cuda(), mel.cuda(), stop.cuda())
model.train_step(*args)
torch.cuda.synchronize()
print("loss", model.losses())
