"""Host arithmetic of the length-group partition: `dtx_plan_length_groups` is what `dtx_step` runs on the row
lengths of a ragged LoRA micro-batch before it touches the device - testable without a GPU."""
import numpy as np
import pytest

from datatunerx_b200 import lib as L


def c128(n, cap):
    return min(cap, max(128, (min(max(int(n), 0), cap) + 127) // 128 * 128))


def check_partition(groups, lens, S_batch):
    rows = [r for g, _ in groups for r in g]
    assert sorted(rows) == list(range(len(lens))), "every row exactly once"
    order = [lens[r] for r in rows]
    if len(groups) > 1:
        assert order == sorted(order, reverse=True), "rows sorted by length, longest first"
    for g, S in groups:
        assert S % 128 == 0 and S <= S_batch
        assert S >= max(c128(lens[r], S_batch) for r in g), "a group is padded to at least its longest row"
        if len(groups) > 1:
            assert len(g) * S >= 256, "every group keeps whole 256-row GEMM tiles"


@pytest.fixture()
def llama7b(lib):
    return L.ModelConfig.llama2_7b()


def test_equal_lengths_stay_one_pass(llama7b):
    assert L.plan_length_groups(llama7b, [512] * 8, 512) == [(list(range(8)), 512)]
    assert L.plan_length_groups(llama7b, [100, 90, 128, 1], 128) == [([0, 1, 2, 3], 128)]
    assert L.plan_length_groups(llama7b, [2048], 2048) == [([0], 2048)]  # a single row has nothing to split


def test_long_tail_is_cut_and_padding_shrinks(llama7b):
    lens = [384, 130, 7, 257, 2048, 600, 512, 900]
    groups = L.plan_length_groups(llama7b, lens, 2048)
    check_partition(groups, lens, 2048)
    assert len(groups) >= 3 and groups[0] == ([4], 2048)
    assert sum(len(g) * S for g, S in groups) < 0.5 * 8 * 2048


def test_partitions_of_the_bench_distribution(llama7b):
    """bench.py --config 7b_varlen: log-normal lengths, 54 % padding in one pass; the plan must cover every row once and never
    pad more than the single pass does."""
    single = planned = real = 0
    n_groups = []
    for step in range(60):
        rng = np.random.default_rng(777 + step)
        lens = np.clip(np.exp(rng.normal(np.log(512.0), 0.6, size=8)), 16, 2048).astype(np.int32)
        S_batch = c128(lens.max(), 2048)
        groups = L.plan_length_groups(llama7b, lens, S_batch)
        check_partition(groups, lens.tolist(), S_batch)
        n_groups.append(len(groups))
        single += 8 * S_batch
        planned += sum(len(g) * S for g, S in groups)
        real += int(lens.sum())
    assert planned <= single and real / planned > 0.7 > real / single, (real / single, real / planned)
    assert 2.0 <= np.mean(n_groups) <= 5.0, np.mean(n_groups)


def test_group_cost_moves_the_cut_and_force_mode_minimises_padding(llama7b):
    lens = [2048, 1500, 1000, 700, 500, 300, 200, 100]
    try:
        L.set_option("varlen_group_cost", 100000)  # a prohibitive fixed cost per group: one pass
        assert len(L.plan_length_groups(llama7b, lens, 2048)) == 1
        L.set_option("varlen_group_cost", 0)
        free = L.plan_length_groups(llama7b, lens, 2048)
        L.set_option("varlen_group_cost", 200)
        default = L.plan_length_groups(llama7b, lens, 2048)
        assert len(free) >= len(default) > 1
        L.set_option("varlen_split", 2)  # parity tests: cut wherever 128-rounded lengths differ (tiny models never split otherwise)
        tiny = L.ModelConfig(vocab=2048, hidden=256, n_layers=2, n_heads=2, ffn=768)
        forced = L.plan_length_groups(tiny, [384, 130, 7, 257], 384)
        check_partition(forced, [384, 130, 7, 257], 384)
        assert len(forced) == 2
        L.set_option("varlen_split", 0)
        assert len(L.plan_length_groups(llama7b, lens, 2048)) == 1
    finally:
        L.set_option("varlen_split", 1)
        L.set_option("varlen_group_cost", 200)


def test_bad_arguments_are_refused(llama7b):
    with pytest.raises(L.DtxError):
        L.plan_length_groups(llama7b, [10, 20], 100)  # padded length must be a multiple of 128


def test_packed_layout(lib):
    """dtx_plan_packed_rows: sequences back to back at 128-rounded lengths (at least one tile each), the default execution of a
    ragged micro-batch."""
    starts, saves = L.plan_packed_rows([384, 130, 7, 257, 0], 384)
    assert starts == [0, 384, 640, 768, 1152, 1280] and saves
    assert all(s % 128 == 0 for s in starts)
    starts, saves = L.plan_packed_rows([100, 100, 100, 100], 128)
    assert starts == [0, 128, 256, 384, 512] and not saves  # as many rows as the padded rectangle: one pass at the padded shape
    starts, saves = L.plan_packed_rows([5000, 2048], 2048)  # lengths are clamped to the batch's padded length
    assert starts == [0, 2048, 4096] and not saves
    rng = np.random.default_rng(1)
    tot_real = tot_packed = tot_padded = 0
    for _ in range(200):  # the bench's distribution: packing keeps less than 12 % padding where the rectangle has 54 %
        lens = np.clip(np.exp(rng.normal(np.log(512.0), 0.6, size=8)), 16, 2048).astype(np.int32)
        S_batch = c128(lens.max(), 2048)
        starts, _ = L.plan_packed_rows(lens, S_batch)
        assert all(b - a >= max(128, int(n)) and (b - a) % 128 == 0 for a, b, n in zip(starts, starts[1:], lens))
        tot_real, tot_packed, tot_padded = tot_real + int(lens.sum()), tot_packed + starts[-1], tot_padded + 8 * S_batch
    assert tot_real / tot_packed > 0.88 and tot_real / tot_padded < 0.5
    with pytest.raises(L.DtxError):
        L.plan_packed_rows([10, 20], 100)
