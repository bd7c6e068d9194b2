"""Model check of the tile handshake that chains consecutive gated layers inside k_flow_tc / between k_layer_tc launches
(parallel-wavenet-vocoder_b200/csrc/pwv_tc.cuh: tiles_ready / publish / cta_of / tiles_in).

Host-side restatement of the protocol's arithmetic, run against an event simulation with RANDOM schedules:
every tile slot of every CTA walks its tile list (layer-major); a tile may load its inputs once the (up to) five
flags of the previous layer it names are set -- its own tile (x[t] rows), the tile(s) holding rows t0-d .. t0+127-d,
and the tile(s) of the previous layer whose x[t-d'] window reads the rows it is about to overwrite (two ping-pong
buffers) -- and it publishes its flag when its output is stored, before it waits for anything else. The simulation
tracks, per row of both buffers, which layer's output the row holds, and asserts
  * RAW: every row a tile reads holds the previous layer's output at the moment of the read,
  * WAR: a tile's store finds the rows holding the layer before the previous one (nobody still needs them),
  * progress: some slot can always move until all tiles are done (no deadlock), for any tile-to-CTA assignment
    (fixed or rotated per layer) and for loads issued arbitrarily early after the flags were seen.
No GPU, no library: this pins the dependency SET; the GPU bit-identity tests pin its implementation."""
import random

import pytest

TM = 128


def tile_deps(t, tiles_per_utt, k, d, d_prev):
    """Tiles of the previous layer (same utterance) that tile k depends on: tiles_ready() in pwv_tc.cuh."""
    t0 = k * TM
    last = tiles_per_utt - 1
    hi, lo = t0 + TM - 1 - d, max(t0 - d, 0)
    k1, k2 = (lo // TM, hi // TM) if hi >= 0 else (k, k)
    k3, k4 = min(k + d_prev // TM, last), min(k + (d_prev + TM - 1) // TM, last)
    return {k, k1, k2, k3, k4}


def assignment(tiles_body, ctas, n_layers, rotate):
    """[layer][cta][slot] -> list of tiles: cta_of / tiles_in of k_flow_tc."""
    rot = tiles_body % ctas if (rotate and tiles_body >= 2 * ctas) else 0
    out = []
    for l in range(n_layers):
        per_cta = []
        for c in range(ctas):
            v = (c + l * rot) % ctas if rot else c
            mine = list(range(v, tiles_body, ctas))
            per_cta.append([mine[0::2], mine[1::2]])
        out.append(per_cta)
    return out


def simulate(n_utt, t, dilations, ctas, rotate, seed):
    rng = random.Random(seed)
    tpu = (t + TM - 1) // TM
    tiles_body = n_utt * tpu
    L = len(dilations)
    assign = assignment(tiles_body, ctas, L, rotate)
    # row versions: buffer b holds for every row the index of the layer whose output it is (-1: the flow's input)
    rows = [[-1] * (n_utt * t), [-9] * (n_utt * t)]
    flags = [[False] * tiles_body for _ in range(L)]
    # every slot: its tile list (layer, tile) and a program counter; states: 0 wait for flags, 1 loaded (inputs read), 2 stored
    slots = []
    for c in range(ctas):
        for s in range(2):
            lst = [(l, k) for l in range(L) for k in assign[l][c][s]]
            if lst:
                slots.append({'list': lst, 'pc': 0, 'state': 0})
    covered = sorted(k for c in range(ctas) for s in range(2) for k in assign[0][c][s])
    assert covered == list(range(tiles_body))                       # every tile has exactly one owner (per layer)
    done = 0
    total = L * tiles_body
    while done < total:
        runnable = []
        for sl in slots:
            if sl['pc'] >= len(sl['list']):
                continue
            l, tile = sl['list'][sl['pc']]
            n, k = divmod(tile, tpu)
            if sl['state'] == 0:
                if l == 0 or all(flags[l - 1][n * tpu + q] for q in tile_deps(t, tpu, k, dilations[l], dilations[l - 1])):
                    runnable.append(sl)
            else:
                runnable.append(sl)
        assert runnable, 'deadlock: %d of %d tiles done' % (done, total)
        sl = rng.choice(runnable)
        l, tile = sl['list'][sl['pc']]
        n, k = divmod(tile, tpu)
        t0, d = k * TM, dilations[l]
        src, dst = rows[l & 1], rows[(l + 1) & 1]
        if sl['state'] == 0:            # flags seen: the loads may land any time from now on -> read now or later (random)
            sl['state'] = 1
            if rng.random() < 0.5:
                continue                # (stay "loading": the read happens in a later step)
        if sl['state'] == 1:
            for r in range(t0, min(t0 + TM, t)):
                assert src[n * t + r] == l - 1, ('RAW x[t]', l, n, k, r, src[n * t + r])
                if r - d >= 0:
                    assert src[n * t + r - d] == l - 1, ('RAW x[t-d]', l, n, k, r - d, src[n * t + r - d])
            sl['state'] = 2
            continue
        # state 2. The slot's NEXT tile may already have its inputs in flight (x[t-d] boxes are refilled as soon as
        # they are converted, long before this tile is stored): if its flags are set, let its read happen first, sometimes.
        if not sl.get('prefetched') and sl['pc'] + 1 < len(sl['list']) and rng.random() < 0.5:
            l2, tile2 = sl['list'][sl['pc'] + 1]
            n2, k2 = divmod(tile2, tpu)
            if l2 == 0 or all(flags[l2 - 1][n2 * tpu + q] for q in tile_deps(t, tpu, k2, dilations[l2], dilations[l2 - 1])):
                src2, d2 = rows[l2 & 1], dilations[l2]
                for r in range(k2 * TM, min(k2 * TM + TM, t)):
                    assert src2[n2 * t + r] == l2 - 1, ('RAW x[t] (prefetch)', l2, n2, k2, r)
                    if r - d2 >= 0:
                        assert src2[n2 * t + r - d2] == l2 - 1, ('RAW x[t-d] (prefetch)', l2, n2, k2, r - d2)
                sl['prefetched'] = True
                continue
        for r in range(t0, min(t0 + TM, t)):
            assert dst[n * t + r] in (l - 2, -9), ('WAR', l, n, k, r, dst[n * t + r])
            dst[n * t + r] = l
        flags[l][tile] = True
        sl['pc'] += 1
        sl['state'] = 2 if sl.pop('prefetched', False) else 0       # (a prefetched tile has read its inputs already)
        done += 1
    assert all(v == L - 1 for v in rows[L & 1])


CASES = [
    # n_utt, T, dilations, CTAs per body, rotate
    (1, 16000, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512], 74, False),     # c1: one or two tiles per CTA
    (2, 4000, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512], 7, True),        # several tiles per slot, rotation
    (3, 1040, [1, 512, 2, 300, 129, 127, 1], 4, True),                  # ragged last tile, d not a multiple of 128, d > T/2
    (2, 1000, [512, 1, 1024, 2000, 3], 5, False),                       # d >= T (no x[t-d] rows at all), descending dilations
    (1, 128, [1, 2, 4], 1, False),                                      # a single tile
    (4, 2000, [256, 1, 256, 1, 64, 640], 3, True),
]


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('seed', [0, 1, 2])
def test_flag_protocol_is_safe_and_live(case, seed):
    n_utt, t, dilations, ctas, rotate = case
    simulate(n_utt, t, dilations, ctas, rotate, seed)


def test_dependency_set_is_minimal_enough_to_matter():
    """Dropping the write-after-read part of the set (k3, k4) must make the simulation fail: the model has teeth."""
    def weak_deps(t, tpu, k, d, d_prev):
        t0 = k * TM
        hi, lo = t0 + TM - 1 - d, max(t0 - d, 0)
        k1, k2 = (lo // TM, hi // TM) if hi >= 0 else (k, k)
        return {k, k1, k2}
    global tile_deps
    saved = tile_deps
    tile_deps = weak_deps
    try:
        failures = 0
        for seed in range(6):
            try:
                simulate(2, 4000, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512], 7, True, seed)
            except AssertionError:
                failures += 1
        assert failures > 0
    finally:
        tile_deps = saved
