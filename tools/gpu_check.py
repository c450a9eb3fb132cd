"""Ad-hoc GPU bring-up check (developer tool, not a test): builds a seeded synthetic workload on the box,
runs the CUDA path through the C ABI and compares with the oracle; prints a first throughput figure."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from helpers import SynthDB, build_fmi, Oracle, make_params as oparams, KoCounters
import kaiju_b200 as kb

def main():
    nprot = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    ncheck = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    nbig = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
    modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["mem"]
    work = "/tmp/kjwork"; os.makedirs(work, exist_ok=True)
    t = time.time(); db = SynthDB(nprot, 1); db.write(work + "/db.faa", work + "/nodes.dmp")
    fmi = build_fmi(work + "/db.faa", work + "/db", threads=os.cpu_count()); print("db+index %.1fs" % (time.time() - t), flush=True)
    orc = Oracle(fmi, work + "/nodes.dmp")
    s1, o1, s2, o2 = db.reads(7, 0, ncheck, 150, True)
    for mode in modes:
        for seg in (True, False):
            P = oparams(mode, seg=seg)
            t = time.time(); tax, best = orc.classify_batch(P, s1, o1, s2, o2); to = time.time() - t
            clf = kb.Classifier(fmi, work + "/nodes.dmp", device=0, params=kb.make_params(mode, seg=seg))
            t = time.time(); gt, gb = clf.classify(s1, o1, s2, o2); tg = time.time() - t
            bad = np.nonzero((gt != tax) | (gb != best))[0]
            print(mode, "seg", seg, "diffs", len(bad), "of", ncheck, "oracle %.2fs gpu %.3fs" % (to, tg), "geom", clf.launch_geometry, "index MB", clf.index_bytes >> 20, flush=True)
            for i in bad[:10]: print("   read", i, "oracle", tax[i], best[i], "gpu", gt[i], gb[i])
            # SE too
            tax1, best1 = orc.classify_batch(P, s1, o1)
            g1, b1 = clf.classify(s1, o1)
            print("   SE diffs", int(((g1 != tax1) | (b1 != best1)).sum()))
            if seg and nbig:
                B1, O1, B2, O2 = db.reads(11, 0, nbig, 150, True)
                for rep in range(3):
                    t = time.time(); gt2, _ = clf.classify(B1, O1, B2, O2); dt = time.time() - t
                    print("   big %d pairs: %.3fs -> %.2f M pairs/s (host buffers, pageable)" % (nbig, dt, nbig / dt / 1e6), flush=True)
                import torch
                d = [torch.from_numpy(x.view(np.uint8) if x.dtype == np.uint8 else x.view(np.int64)).cuda() for x in (B1, O1, B2, O2)]
                dt_ = torch.zeros(nbig, dtype=torch.int64, device="cuda"); db_ = torch.zeros(nbig, dtype=torch.int32, device="cuda")
                for rep in range(3):
                    clf.classify_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), nbig, dt_.data_ptr(), db_.data_ptr(), 150, 150, None)
                    torch.cuda.synchronize(); ms = clf.last_kernel_ms
                    print("   device-resident kernel: %.2f ms -> %.2f M pairs/s" % (ms, nbig / ms / 1e3), flush=True)
                assert (dt_.cpu().numpy().view(np.uint64) == gt2).all()
            clf.close()

if __name__ == "__main__":
    main()
