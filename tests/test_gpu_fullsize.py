"""Size-independent properties at BASELINE.json's full size (1 000 000 series x 256, ComprehensiveFCParameters):
the oracle cannot run there, so the GPU result is checked (a) against float64 torch reductions for the columns
that have closed forms, on every one of the 1 M rows, (b) for batch invariance -- any row of the big run equals
the same series extracted in a small batch, bit for bit, (c) for run-to-run determinism."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_one_million_series_properties():
    torch = pytest.importorskip("torch")
    from tsfresh_b200 import _lib
    from tsfresh_b200.plan import Plan
    from tsfresh_b200.settings import ComprehensiveFCParameters

    S, L = 1_000_000, 256
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    gen = torch.Generator(device=dev)
    gen.manual_seed(44)
    values = torch.randn((S, L), generator=gen, device=dev, dtype=torch.float32)
    plan = Plan(ComprehensiveFCParameters())
    F = plan.n_cols
    ctx = _lib.Context(0, stream=stream.cuda_stream)
    dp = _lib.DevicePlan(ctx, plan)
    out = torch.empty((S, F), device=dev, dtype=torch.float64)
    dp.extract_dense_device(values.data_ptr(), S, L, out.data_ptr())
    torch.cuda.synchronize()
    col = {s: i for i, s in enumerate(plan.suffixes)}
    x = values.double()

    def close(name, ref, rtol=1e-9):
        got = out[:, col[name]]
        assert torch.allclose(got, ref, rtol=rtol, atol=1e-12, equal_nan=True), name

    assert bool((out[:, col["length"]] == L).all())
    close("sum_values", x.sum(1))
    close("mean", x.mean(1))
    close("abs_energy", (x * x).sum(1))
    close("variance", x.var(1, unbiased=False))
    close("standard_deviation", x.std(1, unbiased=False))
    assert bool((out[:, col["maximum"]] == x.max(1).values).all())
    assert bool((out[:, col["minimum"]] == x.min(1).values).all())
    srt = x.sort(1).values
    assert bool((out[:, col["median"]] == 0.5 * (srt[:, L // 2 - 1] + srt[:, L // 2])).all())
    close("quantile__q_0.9", torch.quantile(x, 0.9, dim=1), rtol=1e-12)
    close("absolute_sum_of_changes", (x[:, 1:] - x[:, :-1]).abs().sum(1))
    assert bool((out[:, col["count_above_mean"]] == (x > x.mean(1, keepdim=True)).sum(1)).all())
    assert bool((out[:, col["first_location_of_maximum"]] == x.argmax(1).double() / L).all())
    spec = torch.fft.rfft(x, dim=1)
    close('fft_coefficient__attr_"real"__coeff_7', spec[:, 7].real, rtol=1e-7)
    close('fft_coefficient__attr_"abs"__coeff_99', spec[:, 99].abs(), rtol=1e-7)
    mu = x.mean(1, keepdim=True)
    xc = x - mu
    ac3 = (xc[:, :-3] * xc[:, 3:]).sum(1) / ((L - 3) * x.var(1, unbiased=False))
    close("autocorrelation__lag_3", ac3, rtol=1e-8)
    # every column is finite or NaN by definition only: the always-NaN column and nothing else is all-NaN
    nan_cols = torch.isnan(out).all(0).nonzero().flatten().tolist()
    assert [plan.suffixes[i] for i in nan_cols] == ["query_similarity_count__query_None__threshold_0.0"]

    # (b) batch invariance, (c) determinism
    idx = torch.tensor([0, 1, 31, 32, 12345, 500_000, 999_999], device=dev)
    small = torch.empty((len(idx), F), device=dev, dtype=torch.float64)
    sub = values[idx].contiguous()
    dp.extract_dense_device(sub.data_ptr(), len(idx), L, small.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(small, nan=-7.0), torch.nan_to_num(out[idx], nan=-7.0))
    checksum = torch.nan_to_num(out, nan=0.0).sum(0)
    out2 = torch.empty_like(out)
    dp.extract_dense_device(values.data_ptr(), S, L, out2.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(out2, nan=0.0).sum(0), checksum)
    assert torch.equal(torch.nan_to_num(out2, nan=-7.0), torch.nan_to_num(out, nan=-7.0))
    dp.close()
    ctx.close()
