"""Batched / streaming front-end for the hot path.

The reference API takes one prompt per call (api/ezaudio.py:101, batch hard-wired to 1 at src/inference.py:67).  A server sees a
stream of independent requests; this module groups the ones that can share a launch sequence (same clip length, step count and
guidance constants -> same captured CUDA graphs), cuts them into batches of at most `max_batch` prompts, assigns the batches to
ranks (prompts are independent units: `ezaudio_b200.shard`), runs them through `EzAudio.generate_audio(list[str], ...)` and hands
the waveforms back in completion order (`stream`) or in request order (`run`).  Per-request seeds are honoured: prompt i of a batch
draws from torch.Generator(seed_i), so a request's audio does not depend on what it was batched with.

Pure host logic: the backend is any object with the reference-shaped `generate_audio(text, length=, guidance_scale=, guidance_rescale=,
ddim_steps=, eta=, random_seed=)`; tests drive it with a stub on CPU."""
from __future__ import annotations

import collections
import dataclasses
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple


@dataclasses.dataclass(frozen=True)
class Request:
    """One text-to-audio request: the arguments of api/ezaudio.py:101-103 (defaults included)."""
    prompt: str
    length: int = 10
    guidance_scale: float = 5
    guidance_rescale: float = 0.75
    ddim_steps: int = 100
    eta: float = 1
    random_seed: Optional[int] = None

    def group_key(self) -> Tuple:
        # "" switches guidance off for the whole call (api/ezaudio.py:109-111): empty prompts only batch with empty prompts
        return (self.length, float(self.guidance_scale or 0.0), float(self.guidance_rescale or 0.0), int(self.ddim_steps), float(self.eta or 0.0),
                self.prompt == "")


@dataclasses.dataclass
class Batch:
    key: Tuple
    tickets: List[int]
    requests: List[Request]


def plan_batches(requests: Sequence[Request], max_batch: int) -> List[Batch]:
    """Stable grouping: requests keep their arrival order inside a group; groups are emitted in order of their first request."""
    if max_batch < 1:
        raise ValueError("max_batch must be >= 1")
    groups: Dict[Tuple, List[int]] = collections.OrderedDict()
    for i, r in enumerate(requests):
        groups.setdefault(r.group_key(), []).append(i)
    out: List[Batch] = []
    for key, idx in groups.items():
        for s in range(0, len(idx), max_batch):
            part = idx[s:s + max_batch]
            out.append(Batch(key, part, [requests[i] for i in part]))
    return out


def batches_of_rank(batches: Sequence[Batch], world: int, rank: int) -> List[Batch]:
    """Whole batches are dealt round-robin: every rank replays the same graph shapes, no collective is needed."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return [b for i, b in enumerate(batches) if i % world == rank]


class BatchingFrontEnd:
    def __init__(self, backend, max_batch: int = 4, world: int = 1, rank: int = 0):
        self.backend, self.max_batch, self.world, self.rank = backend, int(max_batch), int(world), int(rank)
        self._queue: List[Request] = []

    def submit(self, prompt: str, **kw) -> int:
        """Queues a request; returns its ticket (index in submission order)."""
        self._queue.append(Request(prompt, **kw))
        return len(self._queue) - 1

    def _run_batch(self, b: Batch):
        r0 = b.requests[0]
        seeds = [r.random_seed for r in b.requests]
        seed_arg = seeds if any(s is not None for s in seeds) else None
        if seed_arg is not None and any(s is None for s in seeds):
            raise ValueError("a batch mixes seeded and unseeded requests: give every request a seed or none")
        sr, wavs = self.backend.generate_audio([r.prompt for r in b.requests], length=r0.length, guidance_scale=r0.guidance_scale,
                                               guidance_rescale=r0.guidance_rescale, ddim_steps=r0.ddim_steps, eta=r0.eta, random_seed=seed_arg)
        if len(wavs) != len(b.requests):
            raise RuntimeError("backend returned a different number of waveforms than prompts")
        return sr, wavs

    def stream(self, requests: Optional[Iterable[Request]] = None) -> Iterator[Tuple[int, int, object]]:
        """Yields (ticket, sample_rate, waveform) batch by batch, as soon as each batch of THIS rank has finished."""
        reqs = list(requests) if requests is not None else self._queue
        if requests is None:
            self._queue = []
        for b in batches_of_rank(plan_batches(reqs, self.max_batch), self.world, self.rank):
            sr, wavs = self._run_batch(b)
            for t, w in zip(b.tickets, wavs):
                yield t, sr, w

    def run(self, requests: Optional[Iterable[Request]] = None) -> List[Optional[Tuple[int, object]]]:
        """All requests of this rank, in request order: result[i] = (sr, waveform), or None for requests served by other ranks."""
        reqs = list(requests) if requests is not None else list(self._queue)
        if requests is None:
            self._queue = []
        out: List[Optional[Tuple[int, object]]] = [None] * len(reqs)
        for t, sr, w in self.stream(reqs):
            out[t] = (sr, w)
        return out
