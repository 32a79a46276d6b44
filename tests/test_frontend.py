"""Batching front-end (ezaudio_b200/frontend.py): host logic only, driven with a stub backend."""
import pytest

from ezaudio_b200.frontend import BatchingFrontEnd, Request, batches_of_rank, plan_batches


class StubBackend:
    """Records calls; 'waveform' of a prompt = (prompt, seed) so that order and seed routing can be checked."""

    def __init__(self):
        self.calls = []

    def generate_audio(self, text, length=10, guidance_scale=5, guidance_rescale=0.75, ddim_steps=100, eta=1, random_seed=None):
        self.calls.append(dict(text=list(text), length=length, gs=guidance_scale, gr=guidance_rescale, steps=ddim_steps, eta=eta, seed=random_seed))
        seeds = random_seed if isinstance(random_seed, (list, tuple)) else [random_seed] * len(text)
        return 24000, [(p, s, length) for p, s in zip(text, seeds)]


def test_grouping_keeps_order_and_caps_batches():
    reqs = [Request(f"p{i}", length=10 if i % 3 else 5, ddim_steps=50) for i in range(11)]
    batches = plan_batches(reqs, max_batch=4)
    seen = [t for b in batches for t in b.tickets]
    assert sorted(seen) == list(range(11)) and all(len(b.tickets) <= 4 for b in batches)
    for b in batches:
        assert len({r.group_key() for r in b.requests}) == 1 and b.tickets == sorted(b.tickets)
    # groups are emitted in order of their first request: request 0 has length 5
    assert batches[0].requests[0].length == 5


def test_empty_prompt_never_shares_a_batch_with_text():
    """'' switches CFG off for the whole call in the reference (api/ezaudio.py:109-111)."""
    batches = plan_batches([Request("a"), Request(""), Request("b"), Request("")], max_batch=8)
    assert [[r.prompt for r in b.requests] for b in batches] == [["a", "b"], ["", ""]]


def test_run_returns_request_order_and_routes_seeds():
    be = StubBackend()
    fe = BatchingFrontEnd(be, max_batch=2)
    tickets = [fe.submit(f"p{i}", ddim_steps=50, random_seed=100 + i, length=10 if i != 2 else 30) for i in range(5)]
    res = fe.run()
    assert tickets == list(range(5))
    assert [r[1][0] for r in res] == [f"p{i}" for i in range(5)]
    assert [r[1][1] for r in res] == [100 + i for i in range(5)]
    assert res[2][1][2] == 30 and all(c["seed"] is not None and len(c["seed"]) == len(c["text"]) for c in be.calls)
    assert all(len(c["text"]) <= 2 for c in be.calls)


def test_mixed_seeding_in_one_batch_is_rejected():
    fe = BatchingFrontEnd(StubBackend(), max_batch=4)
    fe.submit("a", random_seed=1)
    fe.submit("b")
    with pytest.raises(ValueError):
        fe.run()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_ranks_partition_the_batches(world):
    reqs = [Request(f"p{i}", ddim_steps=50) for i in range(13)]
    batches = plan_batches(reqs, 4)
    parts = [batches_of_rank(batches, world, r) for r in range(world)]
    got = sorted(t for p in parts for b in p for t in b.tickets)
    assert got == list(range(13))
    outs = [BatchingFrontEnd(StubBackend(), 4, world, r).run(reqs) for r in range(world)]
    for i in range(13):
        assert sum(o[i] is not None for o in outs) == 1   # every request is served by exactly one rank


def test_stream_yields_batch_by_batch():
    be = StubBackend()
    fe = BatchingFrontEnd(be, max_batch=3)
    it = fe.stream([Request(f"p{i}", ddim_steps=50) for i in range(7)])
    first = next(it)
    assert first[0] == 0 and len(be.calls) == 1   # nothing beyond the first batch has run yet
    rest = list(it)
    assert [t for t, _, _ in [first] + rest] == list(range(7)) and len(be.calls) == 3
