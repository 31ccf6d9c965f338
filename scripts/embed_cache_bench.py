"""Host-side measurement of the embedding-cache read path: the reference's per-clip compressed npz
(np.load(f)["v"], cvap/data/audioset_cls.py:337) against one packed shard (vipant_b200.embed_cache).  CPU only.
    python scripts/embed_cache_bench.py [n_items] [rows_per_item] > profiles/rNN_embed_cache.json"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vipant_b200 import embed_cache as ec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4875
k = int(sys.argv[2]) if len(sys.argv) > 2 else 5
D = 512
rng = np.random.default_rng(0)
with tempfile.TemporaryDirectory() as tmp:
    root = os.path.join(tmp, "npz"); os.makedirs(root)
    names = [f"clip_{i:06d}" for i in range(n)]
    t0 = time.perf_counter()
    for name in names:
        np.savez_compressed(os.path.join(root, name), v=rng.standard_normal((k, D)).astype(np.float32))
    t_write = time.perf_counter() - t0
    npz_bytes = sum(os.path.getsize(os.path.join(root, f)) for f in os.listdir(root))
    t0 = time.perf_counter(); info = ec.pack_npz_dir(root, os.path.join(tmp, "all.vpae")); t_pack = time.perf_counter() - t0
    t0 = time.perf_counter(); info16 = ec.pack_npz_dir(root, os.path.join(tmp, "all16.vpae"), dtype=ec.DTYPE_BF16); t_pack16 = time.perf_counter() - t0
    order = [names[i] for i in rng.permutation(n)]
    def ref_read():
        return np.concatenate([np.load(os.path.join(root, nm + ".npz"))["v"] for nm in order])
    def shard_read(path):
        sh = ec.EmbeddingShard(path)            # includes opening + parsing the index
        return sh.gather(order)[0]
    def best(fn, reps=3):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
        return min(ts), r
    t_ref, a = best(ref_read)
    t_sh, b = best(lambda: shard_read(os.path.join(tmp, "all.vpae")))
    t_sh16, c = best(lambda: shard_read(os.path.join(tmp, "all16.vpae")))
    assert np.array_equal(a, b)
    payload = n * k * D * 4
    print(json.dumps({
        "items": n, "rows_per_item": k, "dim": D, "payload_mb": payload / 1e6, "cpu_count": os.cpu_count(),
        "npz_dir": {"bytes_on_disk_mb": npz_bytes / 1e6, "read_all_s": t_ref, "items_per_s": n / t_ref, "mb_per_s": payload / t_ref / 1e6,
                    "write_s": t_write},
        "shard_fp32": {"bytes_on_disk_mb": info["bytes"] / 1e6, "pack_s": t_pack, "open_and_gather_s": t_sh, "items_per_s": n / t_sh,
                       "mb_per_s": payload / t_sh / 1e6, "speedup_vs_npz": t_ref / t_sh, "bit_identical": True},
        "shard_bf16": {"bytes_on_disk_mb": info16["bytes"] / 1e6, "pack_s": t_pack16, "open_and_gather_s": t_sh16,
                       "items_per_s": n / t_sh16, "speedup_vs_npz": t_ref / t_sh16},
        "note": "page cache warm for both (best of 3); random item order; one process",
    }, indent=1))
