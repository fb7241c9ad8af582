import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libcxxdes_b200 as L
p = L.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0)
ens = L.Ensemble(p, 1 << 16, seed=1)
ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
ens.run().sync(); full = ens.last_run_ms()
ens.reset()
tot = 0.0
for k in range(10):
    ens.run_for(4_000_000); ens.sync(); tot += ens.last_run_ms()
print(f"aloha 65536 replications: run() {full:.1f} ms; 10 run_for pieces {tot:.1f} ms in total")
ens.close()
