import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from atomorph_b200 import engine as eng, scenes
import torch
e = eng.Engine(0, seed=1, motion=eng.SPLINE, fading=eng.PERLIN, feather=2, fluid=10, threads=0, cycle_length=1000)
e.load_images(scenes.square_to_disc(1024))
e.step(8)
e.swap_rounds(512)
out = torch.empty((4, 1024, 1024), dtype=torch.int32, device="cuda:0")
times = np.array([f / 64.0 for f in range(4)])
e.render_into(times, out.data_ptr(), True)
e.sync()
