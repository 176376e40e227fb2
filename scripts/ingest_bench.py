"""Host-side ingest throughput: EventBatch wire bytes -> columns -> graph, against (a) the protobuf runtime's parse
and (b) the per-event loader graph.graph_from_events.  CPU only (the device CSR stage is timed by bench.py).

    python scripts/ingest_bench.py [replicas]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402

from nerrf_b200 import graph as G, ingest, trace_sim  # noqa: E402


def best(f, reps=3):
    out, dt = None, 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); out = f(); dt = min(dt, time.perf_counter() - t0)
    return out, dt


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    ev = G.replicate_events(trace_sim.lockbit_trace(n_files=45, seed=0), k)
    wire = ingest.encode_event_batch(ev)
    n = len(ev)
    print(f"{n} events, {len(wire) / 1e6:.1f} MB on the wire")
    cols, t_dec = best(lambda: ingest.decode_event_batch(wire))
    print(f"decode_event_batch      : {t_dec * 1e3:8.1f} ms  {n / t_dec / 1e6:6.2f} M events/s  {len(wire) / t_dec / 1e9:5.2f} GB/s")
    g, t_g = best(lambda: ingest.graph_from_columns(cols))
    print(f"graph_from_columns      : {t_g * 1e3:8.1f} ms  {n / t_g / 1e6:6.2f} M events/s  ({g.num_nodes} nodes, {g.num_edges} edges)")
    try:
        from proto_util import classes
        _, Batch = classes()
        b, t_pb = best(lambda: Batch.FromString(wire))
        print(f"protobuf runtime parse  : {t_pb * 1e3:8.1f} ms  {n / t_pb / 1e6:6.2f} M events/s   (objects, not columns)")
        evs, t_ev = best(lambda: G.events_from_event_batch(b), reps=1)
        g2, t_loop = best(lambda: G.graph_from_events(evs), reps=1)
        print(f"events_from_event_batch : {t_ev * 1e3:8.1f} ms")
        print(f"graph_from_events (loop): {t_loop * 1e3:8.1f} ms  {n / t_loop / 1e6:6.2f} M events/s")
        same = np.array_equal(g.rowptr, g2.rowptr) and np.array_equal(g.col, g2.col) and np.array_equal(g.x, g2.x)
        print(f"same graph: {same};  wire -> graph speed-up {(t_pb + t_ev + t_loop) / (t_dec + t_g):.1f}x")
    except ImportError as e:
        print("protobuf runtime not available:", e)


if __name__ == "__main__":
    main()
