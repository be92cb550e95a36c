"""Frame-of-reference page decode on the device: throughput of sr_pages_decode on SSB-shaped key columns.

    python tools/page_decode_bench.py [--rows 200000000] [--page-rows 65536]

Encodes synthetic lineorder key columns with the (reference-pinned) oracle encoder on the host, then times the decode of the
resident pages (device memory) and of pages in page-locked host memory (the decode is the PCIe transfer), and checks a
checksum of every decoded column against the raw values.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402  (encoder only: test infrastructure building the input)
from starrocks_b200 import abi, gpu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=50_000_000)
    ap.add_argument("--page-rows", type=int, default=65_536)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--repeat", type=int, default=8, help="decode the encoded pages this many times over into one column (amortises the per-call costs without more host encoding)")
    args = ap.parse_args()
    n = args.rows
    rng = np.random.default_rng(1)
    cols = {"lo_custkey (22 bits)": rng.integers(1, 3_000_000, n, dtype=np.int32), "lo_suppkey (18 bits)": rng.integers(1, 200_000, n, dtype=np.int32),
            "lo_orderdate (sorted runs)": np.sort(rng.integers(19920101, 19981231, (n + 4095) // 4096 * 4096, dtype=np.int32).reshape(-1, 4096), axis=1).reshape(-1)[:n].copy(),
            "lo_revenue int64 (24 bits)": rng.integers(1, 10_000_000, n, dtype=np.int64)}
    ctx = gpu.Context(0)
    dec = gpu.PageDecoder(ctx)
    res = []
    for name, v in cols.items():
        t0 = time.perf_counter()
        pages = [oracle.for_encode(v[lo:lo + args.page_rows]) for lo in range(0, n, args.page_rows)]
        enc_s = time.perf_counter() - t0
        total = sum(len(p) for p in pages)
        typ = abi.TYPE_INT if v.dtype == np.int32 else abi.TYPE_BIGINT
        R = args.repeat
        out = torch.empty(n * R, dtype=torch.int32 if v.dtype == np.int32 else torch.int64, device="cuda")
        blob = torch.from_numpy(np.concatenate([np.pad(p, (0, (-len(p)) % 16)) for p in pages]))
        offs = np.cumsum([0] + [len(p) + (-len(p)) % 16 for p in pages])
        dblob = blob.cuda()
        pblob = blob.pin_memory()
        line = {"column": name, "rows": n * R, "pages": len(pages) * R, "page_bytes": total * R, "bits_per_value": 8.0 * total / n}
        for label, base, mem in (("device", dblob.data_ptr(), abi.MEM_DEVICE), ("pinned_host", pblob.data_ptr(), abi.MEM_HOST_PINNED)):
            views = [(base + int(offs[k]), len(p)) for k, p in enumerate(pages)] * R
            best = 1e9
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                rows = dec.decode(abi.PAGE_FOR, typ, views, out.data_ptr(), n * R, mem=mem)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            assert rows == n * R
            ok = bool((out.view(R, n).cpu().numpy() == v[None, :]).all())
            line[label] = {"ms": best, "values_per_s": n * R / best * 1e3, "page_gbs": total * R / best / 1e6, "output_gbs": n * R * v.dtype.itemsize / best / 1e6,
                           "bit_exact": ok}
        line["host_encode_s"] = enc_s
        res.append(line)
    dec.close()
    print(json.dumps({"workload": "frame-of-reference page decode (sr_pages_decode)", "page_rows": args.page_rows, "columns": res}))


if __name__ == "__main__":
    main()
