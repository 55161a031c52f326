"""Host-side pieces of bench.py that need no GPU: the algorithmic-byte formula of SURVEY.md 8(d) (the numerator of
`roofline.achieved`), the workload -> sharding decision and the kernel named in the roofline object."""
import pytest

import bench
import stp_scenes as S


def test_algorithmic_bytes_reproduce_the_worked_examples_of_the_survey():
    """'C2 R=6e6 -> 2.89 GB/frame', 'C3 R=24e6 -> 13.8 GB' (M=16, V=P, 1080p).  bench.py leaves out the
    survey's `zero = (108+12M)P` term -- the zero-fill of the gradient tensors, which this library does not perform (every
    dL_* array is a pure output) -- so its totals, and with them `roofline.achieved`, are the smaller, conservative ones."""
    zero = lambda P: (108 + 12 * 16) * P  # noqa: E731
    N, T = 1920 * 1080, 120 * 68
    c2 = bench.algorithmic_bytes(1_000_000, 1_000_000, 6_000_000, N, T, 16, 0, 0)
    c3 = bench.algorithmic_bytes(4_000_000, 4_000_000, 24_000_000, N, T, 16, 1, 1)
    assert (sum(c2.values()) + zero(1_000_000)) / 1e9 == pytest.approx(2.89, abs=0.01)
    assert (sum(c3.values()) + zero(4_000_000)) / 1e9 == pytest.approx(13.8, abs=0.05)
    # closed forms of the same table: fwd = (367+48c)P + (200+48s)R + 20N, bwd = 995P + (40+48s)R + (20+12s)N  (+ O(T))
    P, R = 4_000_000, 24_000_000
    fwd = c3["Preprocess"] + c3["Duplicate"] + c3["Sort"] + c3["Render"]
    bwd = c3["RenderBackward"] + c3["PreprocessBackward"]
    assert fwd == pytest.approx((367 + 48) * P + 248 * R + 20 * N, rel=1e-3)
    assert bwd + zero(P) == pytest.approx(995 * P + 88 * R + 32 * N, rel=1e-3)


def test_sort_stage_reports_its_own_traffic_next_to_the_radix_formula():
    R, T = 24_000_000, 8160
    survey = bench.algorithmic_bytes(4_000_000, 4_000_000, R, 1920 * 1080, T, 16, 1, 1)["Sort"]
    assert survey == 6 * 24 * R + 16 * R + 16 * T  # six radix passes + ranges
    assert bench.own_sort_bytes(R, T, False) == 20 * R + 16 * T
    assert bench.own_sort_bytes(R, T, True) == (20 + 84 + 80) * R + 16 * T


def test_workload_table_and_sharding_decision():
    for name, (scene, overrides, _desc) in bench.WORKLOADS.items():
        assert scene in S.CONFIGS, name
        S.default_settings_dict(**overrides)  # every override is a valid settings key
    assert bench.WORKLOADS["C3b"][0] == "C3"  # the default workload is the 4M configuration the target is quoted on
    assert not bench.bands_mode_of("C3b", "auto", 1)  # one GPU: nothing to shard
    assert bench.bands_mode_of("C3b", "auto", 8) and bench.bands_mode_of("C4", "auto", 2)
    assert not bench.bands_mode_of("C2", "auto", 8) and not bench.bands_mode_of("C5", "auto", 8)  # views
    assert bench.bands_mode_of("C2", "bands", 2) and not bench.bands_mode_of("C4", "views", 2)


def test_roofline_names_the_kernel_of_the_settings():
    preset = S.default_settings_dict(**S.STOPTHEPOP_PRESET)
    assert bench.kernel_name("Render", preset) == "render_hier_kernel<4,8,1,0>"
    assert bench.kernel_name("RenderBackward", preset).startswith("blend_replay_bwd_kernel<1,0>")
    assert bench.kernel_name("Render", S.default_settings_dict()) == "render_global_fwd_kernel"
    assert bench.kernel_name("Render", S.default_settings_dict(sort_mode=2, per_pixel=16)) == "render_kbuffer_kernel<16,fwd>"
    assert bench.kernel_name("Sort", preset).startswith("tile_sort_small_kernel")
