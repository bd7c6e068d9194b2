"""Model check of the mbarrier protocol INSIDE one tile slot of k_layer_h
(parallel-wavenet-vocoder_b200/csrc/pwv_tc2.cuh): producer warp, MMA issuer, worker warps.

A host-side restatement of who waits for / arrives on which barrier in which order, run as an event simulation under
RANDOM schedules (asynchronous completions -- TMA loads, tensor-pipe commits -- fire at arbitrary later times, the
tensor pipe in issue order). A wait is modelled exactly as `mbarrier.try_wait.parity`: it passes when the barrier's
current phase parity differs from the parity asked for. The simulation asserts

  * no parity aliasing: whenever a wait for completion k passes, the barrier has completed exactly k + 1 phases (not
    k - 1: premature; and a waiter left two phases behind blocks forever, which shows up as a deadlock);
  * progress: until every agent has finished, something can always move;
  * the data hazards the protocol exists for: a box is reloaded only after every warp copied it (or the tcgen05.cp
    that reads it retired), the A columns take the next tile only when the GEMMs that read them retired and the warp
    has its x[t] in registers, GEMM1 of the next tile overwrites the accumulators only after every warp read D2, the
    staging boxes are written only after their previous content was consumed and the previous store read them.

Modes: 'legacy' (copy after the read-out: round 2's first form, bf16 with double_a = 0 / f16x3 with z_in_d = 0),
'early_db' (bf16: A double-buffered, copy while GEMM2 runs), 'early_zd' (f16x3: z in the accumulator columns),
'cp' (operands by tcgen05.cp from the MMA issuer), each with and without GEMM2 (the flow's last layer has none),
and 'cp_workers_wait_x' -- the first cp version, whose workers waited on x_full in the gate prologue: the barrier
can be two phases ahead by then, the parity wait aliases, the kernel hung on the GPU. The model must find that.
No GPU, no library: this pins the ORDER of the hand-offs; the GPU bit-identity tests pin their implementation."""
import random

import pytest

W = 2     # worker warps per slot in the model (8 in the kernel: the barriers count warps, any W >= 2 shows the races)


class Hazard(AssertionError):
    pass


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, 0, 0

    def arrive(self):
        self.pending += 1
        if self.pending == self.count:
            self.pending = 0
            self.phase += 1

    def passes(self, k):              # try_wait.parity(k & 1)
        return (self.phase & 1) != (k & 1)


class Sim:
    def __init__(self, tiles, mode, last, rng):
        self.tiles, self.mode, self.last, self.rng = tiles, mode, last, rng
        self.cp = mode.startswith('cp')
        self.early = mode in ('early_db', 'early_zd') and not last
        names = {'x_full': 1, 'y_full': 1, 'ax': W, 'ay': W, 'd1': 1, 'za': W, 'zb': W, 'd2': 1, 'out': W, 'a_free': W, 'boxes_free': 1}
        self.b = {k: Bar(k, c) for k, c in names.items()}
        self.async_tma = []           # pending TMA completions: (bar, on_done)
        self.async_tc = []            # tensor pipe, retires in issue order: (bar, on_done)
        none = lambda: [-1] * W
        self.copyx, self.copyy, self.gate, self.xread, self.e2, self.staged = none(), none(), none(), none(), none(), none()
        self.gemm1 = self.gemm2 = self.cpdone = self.xload = self.yload = self.store = -1
        self.log = []

    # ---- agents: generators yielding ('wait', bar, k) or ('do', callable)
    def producer(self):
        b, T = self.b, self.tiles
        copied, xcopied = ('boxes_free', 'boxes_free') if self.cp else ('ay', 'ax')

        def issue_x(j):
            def go():
                if self.cp:
                    self.check(self.cpdone >= j - 1, 'X boxes reloaded before tcgen05.cp(%d) retired' % (j - 1))
                else:
                    self.check(min(self.copyx) >= j - 1, 'X boxes reloaded before every warp copied tile %d' % (j - 1))
                self.async_tma.append((b['x_full'], lambda: setattr(self, 'xload', j)))
            return ('do', go)

        def issue_y(j):
            def go():
                if self.cp:
                    self.check(self.cpdone >= j - 1, 'Y boxes reloaded before tcgen05.cp(%d) retired' % (j - 1))
                else:
                    self.check(min(self.copyy) >= j - 1, 'Y boxes reloaded before every warp copied tile %d' % (j - 1))
                self.check(self.store >= j - 2 and (j < 2 or min(self.staged) >= j - 2), 'Y boxes reloaded under the staged output')
                self.async_tma.append((b['y_full'], lambda: setattr(self, 'yload', j)))
            return ('do', go)

        if T > 0:
            yield issue_x(0)
            yield issue_y(0)
        if T > 1:
            yield ('wait', copied, 0)
            yield issue_x(1)
            yield issue_y(1)
        for j in range(T):
            if j + 1 < T:
                yield ('wait', xcopied, j + 1)
                if j + 2 < T:
                    yield issue_x(j + 2)
            yield ('wait', 'out', j)

            def store(j=j):
                self.check(min(self.staged) >= j, 'store of tile %d before every warp staged it' % j)
                self.store = j
            yield ('do', store)
            if j + 2 < T:
                yield issue_y(j + 2)
            elif j + 1 < T:
                yield ('do', b['y_full'].arrive)

    def mma(self):
        b, T = self.b, self.tiles
        for j in range(T):
            if self.cp:
                if j > 0:
                    yield ('wait', 'a_free', j - 1)
                yield ('wait', 'x_full', j)
                yield ('wait', 'y_full', j)

                def cp(j=j):
                    self.check(self.gemm1 >= j - 1 and (self.last or self.gemm2 >= j - 1), 'tcgen05.cp(%d) into A columns a GEMM still reads' % j)
                    self.async_tc.append((b['boxes_free'], lambda: setattr(self, 'cpdone', j)))
                yield ('do', cp)
            else:
                yield ('wait', 'ax', j)
                yield ('wait', 'ay', j)

            def g1(j=j):
                if not self.cp:
                    self.check(min(self.copyx) >= j and min(self.copyy) >= j, 'GEMM1(%d) before its operand was copied' % j)
                if self.last:
                    self.check(min(self.gate) >= j - 1, 'GEMM1(%d) overwrites accumulators the gate still reads' % j)
                else:
                    self.check(min(self.e2) >= j - 1, 'GEMM1(%d) overwrites D2 of tile %d before every warp read it' % (j, j - 1))
                self.async_tc.append((b['d1'], lambda: setattr(self, 'gemm1', j)))
            yield ('do', g1)
            if self.last:
                continue
            yield ('wait', 'za', j)
            yield ('wait', 'zb', j)

            def g2(j=j):
                self.check(min(self.gate) >= j, 'GEMM2(%d) before z was written' % j)
                self.async_tc.append((b['d2'], lambda: setattr(self, 'gemm2', j)))
            yield ('do', g2)

    def worker(self, w):
        b, T, mode = self.b, self.tiles, self.mode

        def a_copy(jn, announce):
            yield ('wait', 'x_full', jn)

            def cx():
                self.check(self.xload >= jn, 'copy of X(%d) before it landed' % jn)
                if self.last:                    # no GEMM2, nothing read back: only GEMM1 of the columns' previous tenant matters
                    prev = jn - 2 if mode == 'early_db' else jn - 1
                    self.check(self.gemm1 >= prev, 'A columns of tile %d still read by GEMM1' % prev)
                elif mode == 'early_db':         # buffer jn & 1 last held tile jn - 2
                    self.check(self.gemm1 >= jn - 2 and self.gemm2 >= jn - 2 and self.e2[w] >= jn - 2, 'A buffer of tile %d still in use' % (jn - 2))
                elif mode == 'early_zd':
                    self.check(self.gemm1 >= jn - 1 and self.xread[w] >= jn - 1, 'A columns of tile %d still in use (zd)' % (jn - 1))
                else:
                    self.check(self.gemm2 >= jn - 1 and self.e2[w] >= jn - 1, 'A columns of tile %d still in use' % (jn - 1))
                self.copyx[w] = jn
            yield ('do', cx)
            yield ('wait', 'y_full', jn)

            def cy():
                self.check(self.yload >= jn, 'copy of Y(%d) before it landed' % jn)
                self.copyy[w] = jn
            yield ('do', cy)
            if announce:
                yield ('do', b['ax'].arrive)
                yield ('do', b['ay'].arrive)

        if T > 0 and not self.cp:
            yield from a_copy(0, True)
        for j in range(T):
            if mode == 'cp_workers_wait_x':
                yield ('wait', 'x_full', j)
            yield ('wait', 'd1', j)

            def gate(j=j):
                self.check(self.gemm1 >= j, 'gate(%d) before GEMM1 retired' % j)
                self.gate[w] = j
            yield ('do', gate)
            early = self.early and j + 1 < T
            if not self.last:
                yield ('do', b['za'].arrive)
                yield ('do', b['zb'].arrive)
                if mode == 'early_zd':
                    yield ('do', lambda j=j: self.xread.__setitem__(w, j))
                if early:
                    yield from a_copy(j + 1, False)
                yield ('wait', 'd2', j)

                def e2(j=j):
                    self.check(self.gemm2 >= j, 'read-out(%d) before GEMM2 retired' % j)
                    self.e2[w] = j
                    self.xread[w] = j
                yield ('do', e2)
            if self.cp:
                yield ('do', b['a_free'].arrive)
                if j + 1 < T:
                    yield ('wait', 'boxes_free', j + 1)
                elif j >= 1:
                    yield ('wait', 'y_full', j + 1)
            elif early:
                yield ('do', b['ax'].arrive)
                yield ('do', b['ay'].arrive)
            elif j + 1 < T:
                yield from a_copy(j + 1, True)
            elif j >= 1:
                yield ('wait', 'y_full', j + 1)

            def stage(j=j):
                if j + 1 < T:
                    consumed = self.cpdone >= j + 1 if self.cp else self.copyy[w] >= j + 1
                    self.check(consumed, 'staging(%d) over boxes of tile %d that were not consumed' % (j, j + 1))
                self.check(self.store >= j - 1, 'staging(%d) while the store of tile %d may still read the boxes' % (j, j - 1))
                self.staged[w] = j
            yield ('do', stage)
            yield ('do', b['out'].arrive)

    def check(self, ok, what):
        if not ok:
            raise Hazard(what)

    def run(self):
        agents = {'producer': self.producer(), 'mma': self.mma()}
        agents.update({'worker%d' % w: self.worker(w) for w in range(W)})
        cur = {}
        for name, g in list(agents.items()):
            cur[name] = next(g, None)
        steps = 0
        while True:
            moves = []
            for name, op in cur.items():
                if op is None:
                    continue
                if op[0] == 'do' or self.b[op[1]].passes(op[2]):
                    moves.append(('agent', name))
            if self.async_tc:
                moves.append(('tc', None))
            moves += [('tma', i) for i in range(len(self.async_tma))]
            if not moves:
                if all(op is None for op in cur.values()):
                    return steps
                stuck = {n: (op[1], op[2], self.b[op[1]].phase) for n, op in cur.items() if op is not None}
                raise Hazard('deadlock: (barrier, completion waited for, phases completed) = %r' % stuck)
            kind, which = self.rng.choice(moves)
            steps += 1
            if kind == 'tc':
                bar, done = self.async_tc.pop(0)
                done()
                bar.arrive()
            elif kind == 'tma':
                bar, done = self.async_tma.pop(which)
                done()
                bar.arrive()
            else:
                op = cur[which]
                if op[0] == 'wait':
                    bar = self.b[op[1]]
                    self.check(bar.phase == op[2] + 1, '%s: wait for completion %d of %s passed with %d phases completed (parity alias)'
                               % (which, op[2], op[1], bar.phase))
                else:
                    op[1]()
                cur[which] = next(agents[which], None)


MODES = ['legacy', 'early_db', 'early_zd', 'cp']


@pytest.mark.parametrize('last', [False, True])
@pytest.mark.parametrize('mode', MODES)
def test_slot_protocol_has_no_alias_deadlock_or_hazard(mode, last):
    for tiles in (1, 2, 3, 4, 7):
        for seed in range(120):
            Sim(tiles, mode, last, random.Random(seed * 31 + tiles)).run()


def test_model_finds_the_cp_hang():
    """The first cp version: workers wait on x_full in the gate prologue. X(j+1) may land before they get there, the
    barrier is then two phases past the one they ask for, the parity wait blocks: measured as a hang on the GPU."""
    found = 0
    for seed in range(200):
        try:
            Sim(4, 'cp_workers_wait_x', False, random.Random(seed)).run()
        except Hazard as e:
            assert 'deadlock' in str(e) or 'parity alias' in str(e)
            found += 1
    assert found > 0


def test_model_finds_a_missing_hand_off():
    """Sanity of the checker itself: announce the next tile BEFORE the read-out (what double_a / z_in_d must not do)
    and the model reports GEMM1 overwriting D2."""
    def broken_run(seed):
        s = Sim(3, 'early_zd', False, random.Random(seed))
        orig = s.worker

        def worker(w):
            ops = list(orig(w))
            # move the announce (the two arrivals on ax / ay after each read-out) in front of the d2 wait
            out, i = [], 0
            while i < len(ops):
                if ops[i][0] == 'wait' and ops[i][1] == 'd2':
                    k = i + 2                                   # wait d2, do e2, then (early) arrive ax, arrive ay
                    if k + 1 < len(ops) and ops[k][0] == 'do' and ops[k][1] == s.b['ax'].arrive:
                        out += [ops[k], ops[k + 1], ops[i], ops[i + 1]]
                        i = k + 2
                        continue
                out.append(ops[i])
                i += 1
            return iter(out)
        s.worker = worker
        s.run()
    found = 0
    for seed in range(200):
        try:
            broken_run(seed)
        except Hazard as e:
            assert 'overwrites D2' in str(e), str(e)
            found += 1
    assert found > 0
