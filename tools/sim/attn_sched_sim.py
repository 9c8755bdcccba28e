#!/usr/bin/env python
"""Developer tool: discrete-event model of the attention kernel's barrier protocol (one MMA issuer with an in-order
tensor pipe, two softmax warpgroups, a single shared S buffer).  Used to compare issue orders before spending GPU
time.  All times in cycles."""
import sys

def simulate(order, n=40, E=520, L=100, LD=140, M=2000, ST=120, qdepth=64):
    """order: list of steps per iteration among 'QK0','QK1','PV0','PV1' with the iteration offset of each,
    e.g. [('QK1',1),('PV0',0),('QK0',2),('PV1',0)].  Returns average period."""
    INF = float('inf')
    s_ready = {}    # (t,j) -> time S_t(j) complete
    s_free = {}     # (t,j) -> time MMA warp may overwrite S (softmax loaded it) as seen by the MMA warp
    p_full = {}     # (t,j) -> time P_t(j) visible to MMA warp
    p_free = {}     # (t,j) -> time PV_t(j) complete (as seen by softmax)
    wg_free = {0: 0.0, 1: 0.0}
    pipe_free = 0.0
    issue_t = 0.0
    last_owner = None
    # prologue: QK0(0), QK1(0), QK0(1)
    prog = [('QK0', 0), ('QK1', 0), ('QK0', 1)]
    for j in range(n):
        for name, off in order:
            jj = j + off
            if name.startswith('QK') and jj <= 1 and not (name == 'QK1' and jj == 1):
                continue
            prog.append((name, jj))
    start_t, end_t = {}, {}
    def sm_start(t, j):
        if (t, j) not in start_t:
            prev_end = sm_end(t, j - 1) if j > 0 else 0.0
            start_t[(t, j)] = max(s_ready[(t, j)] + L, prev_end)
        return start_t[(t, j)]
    def sm_end(t, j):
        if (t, j) not in end_t:
            st = sm_start(t, j)
            e = st + LD + M
            if j > 0:
                if (t, j - 1) not in p_free:
                    raise RuntimeError(f"deadlock: softmax {t},{j} needs PV{t}({j-1}) which is not issued yet")
                e = max(e, p_free[(t, j - 1)])
            end_t[(t, j)] = e + ST
        return end_t[(t, j)]
    for name, j in prog:
        t = int(name[2])
        if name.startswith('QK'):
            dep = 0.0
            if last_owner is not None:
                dep = sm_start(*last_owner) + LD + L
            start_issue = max(issue_t, dep)
            start = max(start_issue, pipe_free)
            pipe_free = start + E
            issue_t = start_issue
            s_ready[(t, j)] = pipe_free + 50
            last_owner = (t, j)
        else:
            dep = sm_end(t, j) + L
            start_issue = max(issue_t, dep)
            start = max(start_issue, pipe_free)
            pipe_free = start + E
            issue_t = start_issue
            p_free[(t, j)] = pipe_free + L
    # steady-state period from S_0 readiness
    ks = sorted(k for k in s_ready if k[0] == 0)
    a, b = ks[len(ks) // 2], ks[-2]
    return (s_ready[b] - s_ready[a]) / (b[1] - a[1])

if __name__ == "__main__":
    orders = {
        "QK1(j+1) PV0(j) QK0(j+2) PV1(j)": [('QK1', 1), ('PV0', 0), ('QK0', 2), ('PV1', 0)],
        "PV0(j) QK1(j+1) PV1(j) QK0(j+2)": [('PV0', 0), ('QK1', 1), ('PV1', 0), ('QK0', 2)],
    }
    for M in (900, 1500, 2000, 2400):
        for name, o in orders.items():
            print(f"M={M:5d} {name:34s} period {simulate(o, M=M):7.0f}")
