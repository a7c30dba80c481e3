import sys, json, time
sys.path.insert(0, "/root/repo")
import bench
from simpleworks_b200.binding import Backend, ConstraintSystem, Marlin, Rng
be = Backend(0)
r, _ = bench.marlin_gpu_run(be, 20, 1)
print("2^20 index_s", r["index_s"], json.dumps(r["index_phases_ms"]))
m = Marlin(be)
cs = ConstraintSystem.builtin("random-sparse", 100000, 1, 2026)
for i in range(2):
    rng = Rng()
    srs = m.generate_universal_srs(100000, 25000, 300000, rng)
    m.profile(True)
    t0 = time.perf_counter(); pk, vk = m.generate_proving_and_verifying_keys(srs, cs); t1 = time.perf_counter()
    print("E index", t1 - t0, json.dumps(m.last_phases()))
    t0 = time.perf_counter(); proof = m.generate_proof(cs, pk, rng); t1 = time.perf_counter()
    print("E prove", t1 - t0, json.dumps(m.last_phases()))
    m.profile(False)
