"""The adversarial-input generator of tests/adversarial.py against the oracle (no GPU needed)."""
import numpy as np

from adversarial import reference_sum, terms_decimate, terms_resample, tune_output


def test_tuned_outputs_sit_a_few_ulps_from_an_integer_on_the_requested_side(port):
    plan = port.resample_plan(160, 147, 1)
    rng = np.random.default_rng(7)
    x = port.lcg_s16(plan.num_in, 777).copy()
    picks = []
    for j in range(6):
        o = 3001 + 487 * j
        terms = terms_resample(plan, o)
        side = 1 if j % 2 == 0 else -1
        v, d = tune_output(rng, terms, x, 1.0, side, tries=24)
        picks.append((o, terms, side))
    y = port.resample_run(plan, 1.0, x, plan.num_out)
    for o, terms, side in picks:
        v = reference_sum(terms, x, 1.0)
        d = v - round(v)
        assert abs(d) < 1e-10 and (d >= 0) == (side > 0)
        # the python restatement of the accumulation order is the oracle's: same truncated sample
        assert int(y[o]) == int(np.trunc(v))


def test_decimator_terms_follow_the_oracle_order(port):
    plan = port.decimate_plan(3, 1)
    x = port.lcg_s16(plan.num_in * 2, 99)
    y = port.decimate_run(plan, 1.0, x, len(x) // 3)
    for i in (50, 200, 511):
        v = reference_sum(terms_decimate(plan, i), x, 1.0)
        assert int(y[i]) == int(np.trunc(max(-32768.0, min(32767.0, v))))
