"""bench.py's roofline bookkeeping (host logic, no GPU): algorithmic bytes per shape tag, family shares, and the choice of
the headline kernel / shape together with its ncu DRAM-traffic figure."""
import bench


def test_algorithmic_bytes_follow_design_md():
    # attention: 4 (fwd) / 7 (bwd) tiles of ws x d fp32 per (window, head) == 4 / 7 x B x N x C x 4 bytes
    assert bench.alg_bytes("window_attn_fwd", (8, 196608, 96, 3, 64)) == 4 * 8 * 196608 * 96 * 4
    assert bench.alg_bytes("window_attn_bwd", (8, 196608, 96, 3, 64)) == 7 * 8 * 196608 * 96 * 4
    assert bench.alg_bytes("linear_wgrad", (1000, 288, 96)) == 1000 * (288 + 96) * 4
    assert bench.alg_bytes("mlp_dgrad_gelu", (1000, 96, 384)) == 1000 * (96 + 2 * 384) * 4
    assert bench.alg_bytes("ln_head_fwd", (1000, 96, 10)) == 1000 * (96 + 10 + 2) * 4
    assert bench.alg_bytes("ln_head_bwd", (1000, 96, 10)) == 1000 * (2 * 96 + 10 + 2) * 4
    assert bench.alg_bytes("layernorm_fwd", (1000, 96, 1)) == 1000 * 96 * 4 * 3
    assert bench.alg_bytes("not_a_kernel", (1,)) is None


def test_headline_is_the_dominant_family_at_the_captured_shape():
    traffic = {"linear_wgrad": 2420385000, "linear_wgrad@shape": [1572864, 288, 96], "window_attn_bwd": 4817975000}
    km = {("linear_wgrad", (1572864, 384, 96)): [0.46] * 10, ("linear_wgrad", (1572864, 288, 96)): [0.36] * 8,
          ("window_attn_bwd", (8, 196608, 96, 3, 64)): [1.7] * 4, ("rel_bias_expand", (1,)): [0.01]}
    kernels, roof = bench.summarize_kernels(km, 100.0, 6550.7, "measured", traffic)
    assert set(kernels) == {"linear_wgrad", "window_attn_bwd"}
    assert abs(kernels["linear_wgrad"]["share_of_step"] - (4.6 + 2.88) / 100.0) < 1e-9
    assert kernels["linear_wgrad"]["largest_shape"] == [1572864, 384, 96]
    # the headline uses the shape the ncu capture was taken at, so that traffic and algorithmic bytes are comparable
    assert roof["kernel"].startswith("hs_linear_wgrad") and "[1572864, 288, 96]" in roof["kernel"]
    assert roof["algorithmic_bytes_per_launch"] == 1572864 * 384 * 4 and roof["traffic"] == 2420385000
    assert abs(roof["achieved"] - 1572864 * 384 * 4 / 0.36e-3 / 1e9) < 1e-6 and roof["launches_timed"] == 8
    assert abs(roof["frac"] - roof["achieved"] / 6550.7) < 1e-12 and roof["bound"] == "hbm"


def test_headline_without_a_matching_capture_reports_null_traffic():
    traffic = {"linear_wgrad": 2420385000, "linear_wgrad@shape": [1572864, 288, 96]}
    km = {("linear_wgrad", (4096, 384, 96)): [0.1] * 3}
    _, roof = bench.summarize_kernels(km, 10.0, 6550.7, "measured", traffic)
    assert roof["traffic"] is None and "[4096, 384, 96]" in roof["kernel"]
    _, roof = bench.summarize_kernels({("window_attn_fwd", (1, 64, 32, 1, 64)): [0.01]}, 10.0, 6550.7, "m", {})
    assert roof["traffic"] is None
    assert bench.summarize_kernels({}, 10.0, 6550.7, "m", {}) == ({}, None)
