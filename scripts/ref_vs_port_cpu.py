"""The CPU arm of bench.py times oracle/restate.py (the port), because /root/reference does not exist on the GPU box.  This
script records, ONCE, in the build container where the reference is mounted, the rate of the REFERENCE'S OWN classes
(quantum/decoder_v2_4.py GNNI.forward under the shim of oracle/ref_loader.py, no_grad, anomaly detection off) beside the
port's on the same inputs and threads, so the two can be seen to agree.

    python scripts/ref_vs_port_cpu.py > profiles/r02_reference_vs_port_cpu.txt
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, restate  # noqa: E402


def main():
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    t = torch.empty(30 << 20, dtype=torch.uint8)      # glibc allocator in its steady state (bench.py does the same)
    t.fill_(1)
    del t
    BS, L, T, reps = 128, 5, 15, 8
    with ref_loader.reference_session("quantum"):
        ns = ref_loader.load_reference("quantum/decoder_v2_4.py",
                                       consts={"BATCH_SIZE": str(BS), "run1": str(BS), "run2": str(BS), "L": str(L)}, seed=7)
        dec = ns.GNNI(T)
        dec.load_state_dict(torch.load(ref_loader.REF_ROOT + "/quantum/new_model/decoder_parameters_epoch3.pkl"))
        batch = next(iter(ns.train_loader))
        rows, cols = int(ns.rows), int(ns.cols)
        ei = batch.edge_index[:, : batch.edge_index.size(1) // BS]
        x = batch.x.reshape(BS, rows + cols)
        w = dec.state_dict()
        with torch.no_grad():
            want = dec(batch)
            got = restate.decode("v2_4", ei, rows, cols, x, w, T=T)
            same = bool(torch.equal(got["prob"].reshape(-1, 1), want))
            t_ref, t_port = [], []
            for _ in range(reps):
                t0 = time.perf_counter()
                dec(batch)
                t_ref.append(time.perf_counter() - t0)
                t0 = time.perf_counter()
                restate.decode("v2_4", ei, rows, cols, x, w, T=T)
                t_port.append(time.perf_counter() - t0)
    t_ref.sort()
    t_port.sort()
    print("decoder_v2_4 GNNI(15).forward, toric L = %d (V = %d, C = %d), batch %d (the reference's BATCH_SIZE), fp64, torch CPU %d threads" %
          (L, rows, cols, BS, threads))
    print("  reference's own classes (shimmed, no_grad): median %.4f s/batch = %.0f syndromes/s" % (t_ref[reps // 2], BS / t_ref[reps // 2]))
    print("  oracle/restate.py (the port bench.py times): median %.4f s/batch = %.0f syndromes/s" % (t_port[reps // 2], BS / t_port[reps // 2]))
    print("  outputs bit-identical: %s;  reference time / port time = %.2f (at the reference's batch size its O(B^2) idx cat-loop and"
          " clones, which the port omits, are in the noise: the port is a fair stand-in for the reference's CPU path)" %
          (same, t_ref[reps // 2] / t_port[reps // 2]))


if __name__ == "__main__":
    main()
