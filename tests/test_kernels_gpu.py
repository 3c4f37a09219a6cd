"""GPU parity tests proper (-m gpu): every sm_100a kernel and the whole training step, through the C ABI,
against torch fp32 references / the CPU oracle.  The checks live in tests/gpu_checks.py."""
import pytest

from tests.conftest import has_gpu

pytestmark = pytest.mark.gpu


def _run(name):
    if not has_gpu():
        pytest.skip("no GPU")
    from tests import gpu_checks
    return gpu_checks.ALL[name]()


@pytest.mark.parametrize("name", ["gemm_nt", "gemm_nt_bn64", "gemm_nt_bn128", "gemm_nn", "gemm_nn_bn64", "gemm_tn",
                                  "gemm_tn_nosplit", "gemm_tn_pair", "gemm_kext", "gemm_ragged", "gemm_large", "gemm_single_cta", "gemm_pair_vs_single",
                                  "gemm_rope_epilogue", "gemm_rope_epilogue_7b", "gemm_swiglu_epilogues"])
def test_gemm(name):
    _run(name)


@pytest.mark.parametrize("name", ["rmsnorm", "rmsnorm_small", "rope", "swiglu", "lora_dropout", "nf4", "nf4_pack", "embedding", "cross_entropy", "adamw"])
def test_hbm_kernels(name):
    _run(name)


@pytest.mark.parametrize("name", ["attn_fwd", "attn_fwd_long", "attn_fwd_rescale", "attn_fwd_odd_tiles", "attn_fwd_exp_fma", "attn_bwd_single_tile", "attn_bwd", "attn_bwd_long",
                                  "attn_gqa", "attn_bwd_rope", "attn_varlen", "attn_window", "attn_bench_shape_s2048", "attn_bench_shape_s4096_gqa"])
def test_attention(name):
    _run(name)


@pytest.mark.parametrize("name", ["trainer_tiny", "trainer_gqa", "trainer_dropout", "trainer_qlora", "trainer_unfused", "trainer_deterministic", "trainer_grad_accum", "trainer_100_steps",
                                  "trainer_varlen", "trainer_varlen_groups", "trainer_window", "eval_rows_force_step", "missing_weight_refused", "layer_7b_shape", "trainer_full", "trainer_full_accum", "trainer_full_gqa",
                                  "worker_end_to_end"])
def test_training_step(name):
    _run(name)
