"""Helper of test_index_fuzz.py (run in a subprocess so that a crash of the parser cannot take pytest
down): opens damaged copies of an index file through the C ABI and prints the status histogram."""
import os, sys, random, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sshash_b200
from sshash_b200 import _lib
src = sys.argv[1]; seed = int(sys.argv[2]); n_iter = int(sys.argv[3]); mode = sys.argv[4]
good = bytearray(open(src, "rb").read())
rnd = random.Random(seed)
import tempfile
tmp = os.path.join(tempfile.gettempdir(), "fuzz_%d.sshash" % os.getpid())
stats = {}
for it in range(n_iter):
    b = bytearray(good)
    if mode == "truncate":
        b = b[: rnd.randrange(0, len(b))]
    else:
        for _ in range(rnd.randrange(1, 4)):
            # favour the structural fields: headers of the nested containers are spread over the file,
            # so flip anywhere, with a bias to the first 64 KB
            pos = rnd.randrange(0, min(len(b), int(os.environ.get("FUZZ_SPAN", "65536")))) if rnd.random() < 0.5 else rnd.randrange(0, len(b))
            b[pos] = rnd.randrange(256) if rnd.random() < 0.5 else (b[pos] ^ (1 << rnd.randrange(8)))
    open(tmp, "wb").write(b)
    t0 = time.time()
    try:
        d = sshash_b200.Dictionary(tmp)
        d.close()
        r = "opened"
    except sshash_b200.SshashGpuError as e:
        r = _lib.STATUS.get(e.status, str(e.status))
    dt = time.time() - t0
    stats[r] = stats.get(r, 0) + 1
    if dt > 5: print("slow", it, dt, r)
os.remove(tmp)
# after the damaged files the same process must still be able to open and query the good one (no sticky CUDA error)
try:
    import numpy as np
    d = sshash_b200.Dictionary(src, max_k=0 if "k63" not in src and "k47" not in src else 63)
    q = d.access_batch(np.arange(0, min(1000, d.num_kmers()), dtype=np.uint64))
    assert (d.lookup_batch(q.reshape(-1)) != np.uint64(2**64 - 1)).all()
    d.close()
except sshash_b200.SshashGpuError as e:
    if "no CUDA device" not in str(e):
        raise
print(stats)
