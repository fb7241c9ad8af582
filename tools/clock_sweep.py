import sys
sys.path.insert(0,'/root/repo')
import libcxxdes_b200 as L
for prec, name in ((L.MILLISECONDS,"ms"),(L.MICROSECONDS,"us"),(L.NANOSECONDS,"ns")):
    ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10000), 1<<20, seed=0x5EED)
    ens.time_unit(1, L.SECONDS).time_precision(1, prec)
    for _ in range(2):
        ens.run().sync()
    acc,_ = ens.reduce()
    print(f"precision {name}: {ens.last_run_ms():.2f} ms, errors {acc.n_errors}, events {acc.n_events}", flush=True)
    ens.close()
