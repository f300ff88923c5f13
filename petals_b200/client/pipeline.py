"""Wavefront execution of work items over a chain of stages: the schedule shared by the training path
(client/sequential_autograd.py: micro-batches through spans) and by chunked prompt ingestion (client/inference_session.py:
prompt chunks through stage streams).

Item *i* enters stage *s+1* while item *i+1* enters stage *s*. Every stage has a :class:`Lane` — a worker thread with an inbox —
and items travel from lane to lane in order, so the order on every stage is deterministic (item order) and S stages working on M
items take M + S - 1 slots: bubble fraction (S - 1) / (M + S - 1). An item carries ``error`` (set when its work raised; later lanes
skip it) and ``detached`` (it left the wavefront and is being completed elsewhere)."""
from __future__ import annotations

import queue
import threading
from typing import List, Optional, Sequence


class Lane(threading.Thread):
    """Worker of one stage of the route. Pulls micro-batches in arrival order — which is micro-batch order, since the previous
    lane emits them in order — runs ``work`` on each and passes it on."""

    _STOP = object()

    def __init__(self, name: str, work, downstream: Optional["Lane"], finished: "queue.Queue"):
        super().__init__(name=name, daemon=True)
        self.inbox: "queue.Queue" = queue.Queue()
        self.work, self.downstream, self.finished = work, downstream, finished

    def run(self) -> None:
        while True:
            mb = self.inbox.get()
            if mb is Lane._STOP:
                if self.downstream is not None:
                    self.downstream.inbox.put(Lane._STOP)
                return
            if not mb.detached and mb.error is None:
                try:
                    self.work(mb)
                except BaseException as e:  # noqa: BLE001 - recorded on the micro-batch, re-raised by the caller
                    mb.error = e
            (self.downstream.inbox if self.downstream is not None else self.finished).put(mb)


def run_wave(micro_batches: Sequence, stage_work: Sequence, threaded: bool) -> None:
    """Push every micro-batch through ``stage_work[0], stage_work[1], ...`` in wavefront order."""
    if not threaded or len(stage_work) * len(micro_batches) == 1:
        # same (stage, micro-batch) order a pipeline would produce, on this thread: diagonal by diagonal
        S, M = len(stage_work), len(micro_batches)
        for diag in range(S + M - 1):
            for s in range(max(0, diag - M + 1), min(S, diag + 1)):
                mb = micro_batches[diag - s]
                if not mb.detached and mb.error is None:
                    try:
                        stage_work[s](mb)
                    except BaseException as e:  # noqa: BLE001
                        mb.error = e
        return
    finished: "queue.Queue" = queue.Queue()
    lanes: List[Lane] = []
    for s in reversed(range(len(stage_work))):
        lanes.insert(0, Lane(f"petals-lane-{s}", stage_work[s], lanes[0] if lanes else None, finished))
    for lane in lanes:
        lane.start()
    for mb in micro_batches:
        lanes[0].inbox.put(mb)
    lanes[0].inbox.put(Lane._STOP)
    for _ in micro_batches:
        finished.get()
    for lane in lanes:
        lane.join()


