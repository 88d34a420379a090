"""Interleaving model of the tcgen05 kernel's mbarrier protocol (csrc/umma_gemm.cu).

mbarrier waits test a single phase-parity bit, so a waiter that runs two phases early or late reads
the wrong phase.  This model runs the five roles (x producer, weight producer, MMA issuer, NDQ dequant
groups of 4 warps) under random interleavings with the kernel's ring sizes and checks (a) no deadlock,
(b) no wait ever passes before the event it stands for has really happened.
Usage: python scripts/pipeline_model.py
"""
import random


def run(SW, A, NDQ, nst, seed):
    rnd = random.Random(seed)

    class Bar:
        def __init__(s, cnt):
            s.init = cnt; s.pend = cnt; s.c = 0

        def arrive(s):
            s.pend -= 1
            if s.pend == 0:
                s.c += 1; s.pend = s.init

        def test(s, parity):
            return (s.c % 2) != parity

    wfull = [Bar(1) for _ in range(SW)]; wempty = [Bar(4) for _ in range(SW)]
    xfull = [Bar(1) for _ in range(A)]; xaempty = [Bar(1) for _ in range(A)]; afull = [Bar(4) for _ in range(A)]
    w_landed = [False] * nst; x_landed = [False] * nst; mma_done = [False] * nst; dq_done = [0] * nst
    err = []

    def xprod():
        for it in range(nst):
            s = it % A; ph = (it // A) & 1
            yield ('wait', xaempty[s], ph ^ 1)
            if it >= A and not mma_done[it - A]: err.append(('x producer early', it))
            def f(s=s, it=it): x_landed[it] = True; xfull[s].arrive()
            yield ('act', f)

    def wprod():
        for it in range(nst):
            s = it % SW; ph = (it // SW) & 1
            yield ('wait', wempty[s], ph ^ 1)
            if it >= SW and dq_done[it - SW] < 4: err.append(('w producer early', it))
            def f(s=s, it=it): w_landed[it] = True; wfull[s].arrive()
            yield ('act', f)

    def mma():
        for it in range(nst):
            s = it % A; ph = (it // A) & 1
            yield ('wait', xfull[s], ph)
            if not x_landed[it]: err.append(('mma x early', it))
            yield ('wait', afull[s], ph)
            if dq_done[it] < 4: err.append(('mma a early', it))
            def f(s=s, it=it): mma_done[it] = True; xaempty[s].arrive()
            yield ('act', f)

    def dq(grp):
        for it in range(nst):
            if it % NDQ != grp: continue
            sw = it % SW; wph = (it // SW) & 1; sl = it % A; aph = (it // A) & 1
            if it >= A:
                yield ('wait', xaempty[sl], aph ^ 1)
                if not mma_done[it - A]: err.append(('dequant slot early', it))
            yield ('wait', wfull[sw], wph)
            if not w_landed[it]: err.append(('dequant weights early', it))
            def f(sw=sw, sl=sl, it=it): wempty[sw].arrive(); dq_done[it] += 1; afull[sl].arrive()
            yield ('act', f)

    threads = [xprod(), wprod(), mma()] + [dq(g) for g in range(NDQ) for _ in range(4)]
    cur = [None] * len(threads); done = [False] * len(threads)
    while not all(done):
        progressed = False
        order = list(range(len(threads))); rnd.shuffle(order)
        for i in order:
            if done[i]: continue
            if cur[i] is None:
                try: cur[i] = next(threads[i])
                except StopIteration: done[i] = True; progressed = True; continue
            op = cur[i]
            if op[0] == 'wait':
                if op[1].test(op[2]): cur[i] = None; progressed = True
            else:
                op[1](); cur[i] = None; progressed = True
            if rnd.random() < 0.7: break
        if not progressed:
            if all(done[i] or (cur[i] is not None and cur[i][0] == 'wait' and not cur[i][1].test(cur[i][2]))
                   for i in range(len(threads))):
                return 'deadlock'
    return 'ok' if not err else str(err[:3])


if __name__ == "__main__":
    # (W stages, XA slots, dequant groups, stages per CTA): the kernel's configurations and a few adversarial ones
    bad = 0
    for cfg in [(8, 4, 4, 60), (12, 6, 4, 80), (16, 7, 4, 90), (24, 7, 4, 120), (14, 7, 4, 43), (4, 4, 4, 50), (9, 5, 5, 70)]:
        res = set(run(*cfg, seed) for seed in range(300))
        print(cfg, res)
        bad += res != {'ok'}
    raise SystemExit(bad)
