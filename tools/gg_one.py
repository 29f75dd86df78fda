import sys, numpy as np
sys.path.insert(0, "/root/repo")
from fastest_lap_b200 import gg, workloads
params = workloads._vehicles()["roberto_lot_kart_2016"]
speeds = np.linspace(50.0, 120.0, 32) / 3.6
r = gg.gg_diagram(params, speeds, n_lat=64)
print(r["ax_max_solved"].mean(), r["ax_min_solved"].mean())
