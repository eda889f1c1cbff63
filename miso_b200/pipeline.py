"""Reads in -> posteriors out, with the setup stage hidden behind the GPU.

The reference does everything of a gene inside one call (``splicing_miso``: matching,
ordering, classes, then the MCMC loop -- ``pysplicing/src/miso.c:758-993``), so its CPU is
never idle.  Here the two halves run on different processors: the plan stage
(``misob200_plan_append``: read <-> isoform matching, draw order, read classes, tile
packing; host threads, optionally the matching on the GPU) and the chain kernels.  Run back to
back the B200 would wait for the host most of the time; ``run_pipelined`` overlaps them:
batches of genes are planned by a background thread (ctypes releases the GIL, the C++ stage
is multi-threaded) while the device runs the previous batch.  A batch of ~12k cfg-3 events
still fills the machine (balanced launch policy, ``csrc/run.cu``).
"""
import queue
import threading

from .batch import Plan


def run_pipelined(read_batches, params, outputs=None, match_device=None, depth=2, on_result=None):
    """read_batches: iterable of objects ``Plan.append`` accepts (``ReadBatch``, a synthetic
    ``Workload`` ...), or of callables returning one (a batch can be loaded lazily by the planner
    thread).  Every batch becomes its own plan, run as soon as it is planned.

    outputs: optional list of pre-allocated output dicts (``Plan.alloc_outputs``), one per batch.
    on_result(i, plan, out): called after batch i has run (e.g. to write its ``.miso`` files);
    the planner keeps working meanwhile.  Returns the list of (plan, out).
    """
    q = queue.Queue(maxsize=max(1, depth))
    stop = threading.Event()

    def planner():
        try:
            for i, rb in enumerate(read_batches):
                if stop.is_set():
                    break
                if callable(rb):
                    rb = rb()
                q.put((i, Plan().append(rb, match_device=match_device), None))
        except BaseException as e:      # surfaced in the consumer
            q.put((-1, None, e))
        q.put(None)

    t = threading.Thread(target=planner, name="misob200-planner", daemon=True)
    t.start()
    done = []
    try:
        while True:
            item = q.get()
            if item is None:
                break
            i, plan, err = item
            if err is not None:
                raise err
            out = outputs[i] if outputs is not None else None
            out = plan.run(params, out)
            done.append((plan, out))
            if on_result is not None:
                on_result(i, plan, out)
    finally:
        stop.set()
        while t.is_alive():             # unblock a planner waiting on a full queue
            try:
                q.get_nowait()
            except queue.Empty:
                pass
            t.join(timeout=0.05)
    return done
