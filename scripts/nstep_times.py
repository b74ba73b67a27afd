import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import emojivoice_b200 as ev
from emojivoice_b200 import synthetic
from emojivoice_b200.config import HIFIGAN_V1, VCTK
model = ev.MatchaTTS(**VCTK.constructor_kwargs(), precision="bf16")
model.load_state_dict(synthetic.matcha_state_dict(VCTK, seed=1234))
x, xl, spk = synthetic.phoneme_batch(32, 60, 90, seed=2000)
x, xl, spk = x.cuda(), xl.cuda(), spk.cuda()
for n in (1, 2, 10, 20):
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = model.synthesise(x, xl, n, 0.667, spk, 0.8); e1.record(); torch.cuda.synchronize()
    print(f"n_timesteps={n}: {e0.elapsed_time(e1):.2f} ms")
