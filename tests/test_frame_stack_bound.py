"""Host model of the frame-stack discipline of cycles_warp (gsn_b200/csrc/count_small.cu): a step takes frames from the top
of a depth-sorted stack until they hold 32 candidates, pushes the (at most one) partially consumed frame back and the
children above it, roots enter at the bottom while fewer than 32 frames are stacked.  The kernel sizes the per-warp stack to
32 * (kmax - 2) frames; this checks the invariant behind that number (every depth holds <= 32 frames) and the sortedness
the reverse-order push relies on, on graphs far denser than molecules."""
import numpy as np
import pytest


def simulate(adj, kmax, nwarps, max_steps):
    n = len(adj)
    gt = lambda v: ~((2 << v) - 1)                           # noqa: E731   ids > v
    ticket, stack, steps, peak = 0, [], 0, 0                 # frame = (depth, cand, X, f0)
    seeds_left = True
    while steps < max_steps:
        if len(stack) < 32 and seeds_left:
            remaining = n - ticket
            take = min(remaining // (4 * nwarps) + 1 if remaining > 0 else 1, 32 - len(stack))
            base, ticket = ticket, ticket + take
            if base >= n:
                seeds_left = False
            roots = []
            for v in range(base, min(base + take, n)):
                cand = adj[v] & gt(v) & ((1 << n) - 1)
                if cand and kmax >= 3:
                    roots.append((0, cand, 1 << v, v))
            stack = roots + stack                            # roots go to the bottom
            if not stack:
                if not seeds_left:
                    break
                continue
        if not stack:
            break
        steps += 1
        depths = [f[0] for f in stack]
        assert depths == sorted(depths), 'stack not sorted by depth'
        window = stack[::-1][:32]                            # deepest first
        taken, pairs, keep = 0, [], None
        for fr in window:
            if len(pairs) >= 32:
                break
            taken += 1
            cands = [b for b in range(n) if fr[1] >> b & 1]
            room = 32 - len(pairs)
            pairs += [(fr, j) for j in cands[:room]]
            if len(cands) > room:
                rest = 0
                for j in cands[room:]:
                    rest |= 1 << j
                keep = (fr[0], rest, fr[2], fr[3])
        del stack[len(stack) - taken:]
        children = []
        for (p, _, X, f0), j in pairs:
            pc = p + 1
            Xc = X | (1 << j)
            ext = adj[j] & ~Xc & gt(f0) & ((1 << n) - 1)
            if pc + 2 < kmax and ext:
                children.append((pc, ext, Xc, f0))
        if keep is not None:
            stack.append(keep)
        stack += children[::-1]                              # lane 0 (deepest parent) ends on top
        peak = max(peak, len(stack))
        per_level = np.bincount([f[0] for f in stack], minlength=kmax) if stack else np.zeros(kmax, int)
        assert per_level.max() <= 32, per_level
        assert len(stack) <= 32 * max(kmax - 2, 1)
    return peak, steps


@pytest.mark.parametrize('n,p,kmax,nwarps', [(24, 0.7, 6, 1), (40, 0.5, 8, 2), (64, 0.9, 5, 8), (64, 0.3, 12, 4), (30, 1.0, 7, 1),
                                             (64, 1.0, 4, 1), (12, 0.5, 3, 1)])
def test_stack_never_exceeds_the_kernel_bound(n, p, kmax, nwarps):
    rng = np.random.default_rng(n * 1000 + kmax)
    a = np.triu(rng.random((n, n)) < p, 1)
    a = a | a.T
    adj = [int(sum(1 << j for j in range(n) if a[i, j])) for i in range(n)]
    peak, steps = simulate(adj, kmax, nwarps, max_steps=4000)
    assert steps > 0 and peak <= 32 * max(kmax - 2, 1)
