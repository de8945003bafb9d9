"""Grid-cooperative sweep with many concurrent sources (CTA groups): full box and R = 80-100 at 256^3, full box at 128^3."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(HERE, "perf_probe3.py")).read().split("thin, thick, dlogtau =")[0])  # imports and run()
thin, thick, dlogtau = p.blackbody_tables(1e5, False, -20.0, 4.0, 20000)
NumTau = 20000
N = 256
p.device_init(N, 64); p.photo_table_to_device(thin, thick)
run(N, 1e4, 16, "f1"); run(N, 1e4, 74, "f1"); run(N, 1e4, 148, "f1"); run(N, 100.0, 74, "f1"); run(N, 100.0, 148, "f1"); run(N, 80.0, 148, "f1")
p.device_close()
N = 128
p.device_init(N, 64); p.photo_table_to_device(thin, thick)
run(N, 1e4, 5, "f0"); run(N, 1e4, 74, "f1")
p.device_close()
