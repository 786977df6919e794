#!/usr/bin/env python
"""Derive and verify the min/max programs used by the sliding median-of-13 kernel.

The kernel produces 4 consecutive medians per step from 16 samples e0..e15
(window of output j = e[j .. j+12]).  The 10 samples e3..e12 are common to all
four windows ("core"); of the sorted core only ranks 3..6 can be the median of
any of the four windows, so a sorting network for 10 inputs is pruned to the
operations that reach those four wires.  Verified with the 0-1 principle.
"""
import itertools, sys

NET10 = [(0,5),(1,6),(2,7),(3,8),(4,9),(0,3),(1,4),(5,8),(6,9),(0,2),(3,6),(7,9),
         (0,1),(2,4),(5,7),(8,9),(1,2),(3,5),(4,6),(7,8),(1,3),(2,5),(4,7),(6,8),
         (2,3),(6,7),(3,4),(5,6),(4,5)]

def sorts(net, n):
    for bits in itertools.product((0,1), repeat=n):
        v = list(bits)
        for a,b in net:
            if v[a] > v[b]: v[a], v[b] = v[b], v[a]
        if any(v[i] > v[i+1] for i in range(n-1)): return False
    return True

def prune(net, n, keep):
    """Return list of (a, b, need_min, need_max) for comparators still needed."""
    live = set(keep)
    out = []
    for a,b in reversed(net):
        need_min, need_max = a in live, b in live
        if need_min or need_max:
            out.append((a,b,need_min,need_max))
            live.add(a); live.add(b)
    out.reverse()
    return out

def check_pruned(prog, n, keep):
    for bits in itertools.product((0,1), repeat=n):
        v = list(bits)
        for a,b,mn,mx in prog:
            lo, hi = min(v[a],v[b]), max(v[a],v[b])
            if mn: v[a] = lo
            if mx: v[b] = hi
        s = sorted(bits)
        if any(v[k] != s[k] for k in keep): return False
    return True

def odd_even_merge(a, b, net):
    """Batcher's odd-even merge of two sorted wire lists; appends comparators to `net` and
    returns the wires in output order."""
    if not a: return list(b)
    if not b: return list(a)
    if len(a) == 1 and len(b) == 1:
        net.append((a[0], b[0]))
        return [a[0], b[0]]
    ev = odd_even_merge(a[0::2], b[0::2], net)
    od = odd_even_merge(a[1::2], b[1::2], net)
    out, rest = [ev[0]], ev[1:]
    k = min(len(od), len(rest))
    for i in range(k):
        net.append((od[i], rest[i]))
        out += [od[i], rest[i]]
    return out + od[k:] + rest[k:]

def merge_6_4_mid4():
    """Ranks 3..6 of sorted six U sorted four (median13x8): pruned merge, checked on all
    sorted 0-1 inputs."""
    net = []
    out = odd_even_merge(list(range(6)), list(range(6, 10)), net)
    live, prog = set(out[3:7]), []
    for p, q in reversed(net):
        mn, mx = p in live, q in live
        if mn or mx:
            prog.append((p, q, mn, mx)); live.add(p); live.add(q)
    prog.reverse()
    for ka in range(7):
        for kb in range(5):
            bits = [0] * (6 - ka) + [1] * ka + [0] * (4 - kb) + [1] * kb
            v = list(bits)
            for p, q, mn, mx in prog:
                lo, hi = min(v[p], v[q]), max(v[p], v[q])
                if mn: v[p] = lo
                if mx: v[q] = hi
            assert [v[w] for w in out[3:7]] == sorted(bits)[3:7]
    name = lambda i: f"a{i}" if i < 6 else f"b{i - 6}"
    print("merge(6,4) -> ranks 3..6:", sum(mn + mx for _, _, mn, mx in prog), "min/max ops; outputs on",
          [name(w) for w in out[3:7]])
    for p, q, mn, mx in prog:
        print(f"  ({name(p)},{name(q)}) {'min' if mn else '   '} {'max' if mx else '   '}")

if __name__ == "__main__":
    merge_6_4_mid4()
    assert sorts(NET10, 10)
    prog = prune(NET10, 10, [3,4,5,6])
    assert check_pruned(prog, 10, [3,4,5,6])
    ops = sum(mn + mx for _,_,mn,mx in prog)
    print("comparators kept", len(prog), "min/max ops", ops)
    for a,b,mn,mx in prog:
        print(f"  ({a},{b}) {'min' if mn else '   '} {'max' if mx else '   '}")
