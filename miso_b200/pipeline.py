"""Reads in -> posteriors out, with the setup stage hidden behind the GPU.

The reference does everything of a gene inside one call (``splicing_miso``: matching,
ordering, classes, then the MCMC loop -- ``pysplicing/src/miso.c:758-993``), so its CPU is
never idle.  Here the halves run on different processors, and run back to back the B200 would
wait for the host most of the time.  ``run_pipelined`` overlaps them, batch by batch:

  host setup (``match_device=None``)      planner thread: ``Plan.append`` of batch i+1 (host threads)
                                          main thread:    chain kernels of batch i

  device setup (``match_device=<gpu>``)   feeder thread:  copies in + match_kernel + order_kernel +
                                                          copies out of batch i+2, enqueued on one of
                                                          the library's three stages (``append_begin``)
                                          finisher thread: read classes + tile packing of batch i+1 on
                                                          the host threads (``append_finish``)
                                          main thread:    chain kernels of batch i
                                          (the setup kernels take the SMs between two chain runs:
                                          ~10 ms per 12k events)

ctypes releases the GIL during the library calls and the C++ stages are multi-threaded.  A batch
of ~12k cfg-3 events still fills the machine (balanced launch policy, ``csrc/run.cu``).
"""
import queue
import threading
import time

from .batch import Plan

_END = object()


def run_pipelined(read_batches, params, outputs=None, match_device=None, depth=2, on_result=None, stats=None,
                  release=True):
    """read_batches: iterable of objects ``Plan.append`` accepts (``ReadBatch``, a synthetic
    ``Workload`` ...), or of callables returning one (a batch can be loaded lazily by the first
    pipeline thread).  Every batch becomes its own plan, run as soon as it is planned.

    outputs: optional list of pre-allocated output dicts (``Plan.alloc_outputs``), one per batch.
    on_result(i, plan, out): called after batch i has run (e.g. to write its ``.miso`` files);
    the setup threads keep working meanwhile.  stats: optional dict, receives per-batch seconds
    (``plan``: setup threads busy, ``wait``: main thread waiting for a plan, ``run``: GPU run).
    release: give a batch's device buffers back once it has run and on_result has seen it (its outputs
    are in host memory by then; summaries / comparisons on the device need release=False).
    Returns the list of (plan, out).
    """
    t_plan, t_wait, t_run = [], [], []
    ready = queue.Queue(maxsize=max(1, depth))          # planned batches -> main thread
    stop = threading.Event()
    threads = []

    def guarded(fn, out_q):
        def body():
            try:
                fn()
            except BaseException as e:      # surfaced in the consumer
                out_q.put((-1, None, e))
            out_q.put(_END)
        t = threading.Thread(target=body, name="misob200-" + fn.__name__, daemon=True)
        threads.append(t)
        t.start()

    def batches():
        for i, rb in enumerate(read_batches):
            if stop.is_set():
                return
            yield i, (rb() if callable(rb) else rb)

    if match_device is None:
        def planner():
            for i, rb in batches():
                t0 = time.perf_counter()
                plan = Plan().append(rb)
                t_plan.append(time.perf_counter() - t0)
                ready.put((i, plan, None))
        guarded(planner, ready)
    else:
        begun = queue.Queue(maxsize=2)                  # (the library has three stages: one more is being finished)

        def feeder():
            for i, rb in batches():
                t0 = time.perf_counter()
                plan = Plan().append_begin(rb, match_device)
                begun.put((i, plan, time.perf_counter() - t0))

        def finisher():
            while True:
                item = begun.get()
                if item is _END:
                    return
                i, plan, err = item
                if i < 0:
                    raise err
                t0 = time.perf_counter()
                plan.append_finish()
                t_plan.append(err + time.perf_counter() - t0)
                ready.put((i, plan, None))
        guarded(feeder, begun)
        guarded(finisher, ready)

    done = []
    try:
        while True:
            t0 = time.perf_counter()
            item = ready.get()
            t1 = time.perf_counter()
            if item is _END:
                break
            i, plan, err = item
            if err is not None:
                raise err
            out = outputs[i] if outputs is not None else None
            out = plan.run(params, out)
            t_wait.append(t1 - t0)
            t_run.append(time.perf_counter() - t1)
            done.append((plan, out))
            if on_result is not None:
                on_result(i, plan, out)
            if release:
                plan.release_device()       # the device state goes back to the library's pool: the next batch reuses it
    finally:
        stop.set()
        while any(t.is_alive() for t in threads):       # unblock threads waiting on a full queue
            for q in ([ready] if match_device is None else [ready, begun]):
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
            time.sleep(0.01)
    if stats is not None:
        stats.update(plan=t_plan, wait=t_wait, run=t_run)
    return done
