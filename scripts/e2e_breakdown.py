"""Where the host-side time of one query goes (run on the GPU box)."""
import statistics
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import motionrag_b200 as m  # noqa: E402
from motionrag_b200 import synthetic  # noqa: E402

n = 1_000_000
st = m.EmbeddingStore(768, n, 0)
synthetic.fill_store(st, n, "clustered", seed=0)
st.set_groups(synthetic.groups(n, 0, "cuda"))
src = torch.randint(0, n, (64,), device="cuda")
q = synthetic.queries_from_rows(st.rows_f32()[src], seed=3)
qh = q.cpu().numpy()
ex = (src // 3).to(torch.int32)
exh = ex.cpu().numpy()
cols = {"video": np.array([f"video_{j // 3:07d}.mp4" for j in range(n)]), "start_sec": np.zeros(n), "end_sec": np.ones(n) * 2}
db = m.RAGDatabase.from_store(st, cols)


def med(fn, reps=300):
    for _ in range(20):
        fn(0)
    ts = []
    for i in range(reps):
        t = time.perf_counter()
        fn(i)
        ts.append(time.perf_counter() - t)
    return statistics.median(ts) * 1e3


def dev_only(i):
    st.search(q[i % 64:i % 64 + 1], 12, exclude_group=ex[i % 64:i % 64 + 1])
    torch.cuda.synchronize()


print("device search + sync            %.4f ms" % med(dev_only))
print("search_host (C, graph)          %.4f ms" % med(lambda i: st.search_host(qh[i % 64:i % 64 + 1], 12, exclude_group=exh[i % 64:i % 64 + 1])))
print("search_host certify             %.4f ms" % med(lambda i: st.search_host(qh[i % 64:i % 64 + 1], 12, exclude_group=exh[i % 64:i % 64 + 1], certify=True)))
import ctypes as C
from motionrag_b200 import _cabi
lib = _cabi.load()
p_ = st._params(12, "l2", "auto", 64, "post", 0)
dist = np.empty((1, 12), np.float32); idx = np.empty((1, 12), np.int64); grp = np.empty((1, 12), np.int32); mg = np.empty(1, np.float32)
p_.out_margin = mg.ctypes.data
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def raw(i):
    lib.mrag_search_host(st._h, qh[i % 64:i % 64 + 1].ctypes.data, 1, C.byref(p_), exh[i % 64:i % 64 + 1].ctypes.data,
                         dist.ctypes.data, idx.ctypes.data, grp.ctypes.data, stream)
print("raw ctypes mrag_search_host     %.4f ms" % med(raw))
wh = [f'video != "video_{int(exh[j]):07d}.mp4"' for j in range(64)]
print("db._search (certified)          %.4f ms" % med(lambda i: db._search(qh[i % 64], "text_embedding", 12, wh[i % 64], 30)))
d0, i0, _ = db._search(qh[0], "text_embedding", 12, wh[0], 30)
print("db._records                     %.4f ms" % med(lambda i: db._records(d0, i0, ["video", "start_sec", "end_sec"])))
print("db._exclusion_ids               %.4f ms" % med(lambda i: db._exclusion_ids(wh[i % 64], 1)))
print("RAGDatabase.text_search         %.4f ms" % med(lambda i: db.text_search(qh[i % 64], top_k=12, where=wh[i % 64], select=["video", "start_sec", "end_sec"])))
tm = []
for i in range(100):
    st.search(q[i % 64:i % 64 + 1], 12, exclude_group=ex[i % 64:i % 64 + 1], timings=tm)
print("events: scan %.4f ms, whole call %.4f ms" % (statistics.median(t[0] for t in tm), statistics.median(t[1] for t in tm)))
print("fp32 rechecks:", db.fp32_rechecks)
