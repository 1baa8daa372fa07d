"""Model of the peer-memory exchange of the trace CG (hybridsbp_b200/csrc/api_p2p.cuh): every rank runs, per iteration,
push A -> wait A -> (consume A) -> push B -> wait B -> (consume B) in stream order; a push writes the rank's payload into a
parity slot of every partner's mailbox and then raises its flag there to the new epoch, a wait blocks until all flags of the own
mailbox have reached the epoch.  The ranks are otherwise unsynchronised.  Checked over random interleavings: no rank ever
blocks forever, and every payload a rank consumes is the one its partner wrote for that very iteration (a slot is never
overwritten before it has been read) -- with the parity double-buffering and, as the header argues, even without it."""
import random

PHASES = ("push_a", "wait_a", "use_a", "push_b", "wait_b", "use_b")


def run(world, iters, seed, double_buffer=True):
    rng = random.Random(seed)
    # mailbox[r][phase][parity][src] = iteration tag of the payload; flags[r][phase][src] = epoch
    mailbox = [{ph: [[None] * world for _ in range(2)] for ph in "ab"} for _ in range(world)]
    flags = [{ph: [0] * world for ph in "ab"} for _ in range(world)]
    pc = [0] * world                      # index into the unrolled program of the rank
    nsteps = iters * len(PHASES)
    while any(p < nsteps for p in pc):
        ready = []
        for r in range(world):
            if pc[r] >= nsteps:
                continue
            it, ph = divmod(pc[r], len(PHASES))
            name = PHASES[ph]
            if name.startswith("wait"):
                if all(f >= it + 1 for f in flags[r][name[-1]]):
                    ready.append(r)
            else:
                ready.append(r)
        assert ready, "deadlock"
        r = rng.choice(ready)
        it, ph = divmod(pc[r], len(PHASES))
        name, epoch = PHASES[ph], it + 1
        par = epoch & 1 if double_buffer else 0
        if name.startswith("push"):
            for dst in range(world):
                mailbox[dst][name[-1]][par][r] = it          # payload first ...
            for dst in range(world):
                flags[dst][name[-1]][r] = epoch              # ... then the flag (fence in between)
        elif name.startswith("use"):
            for src in range(world):
                assert mailbox[r][name[-1]][par][src] == it, (r, name, it, src, mailbox[r][name[-1]][par][src])
        pc[r] += 1


def test_no_slot_is_overwritten_before_it_is_consumed():
    for world in (2, 3, 8):
        for seed in range(40):
            run(world, 12, seed)


def test_the_two_phases_order_the_reuse_even_without_double_buffering():
    for world in (2, 4, 8):
        for seed in range(40):
            run(world, 12, 1000 + seed, double_buffer=False)
