"""A model of fit's batched loop (csrc/mmsig_api.cu: mmctm_run_iterations / mmctm_iterate_async), in plain Python.

Beyond the tenth iteration the host enqueues `nb` iterations at a time.  Per enqueued iteration it swaps its pointers
(lam <-> lam_prev, sumtheta <-> sumtheta_alt) BEFORE launching, the device evaluates the stopping rule at the end of an
iteration, and once the rule has fired every later kernel of the batch returns at once -- except that the theta pass of
the NEXT iteration may already have started (it overlaps the side stream that evaluates the rule) and runs to completion.
After the batch the host undoes the swaps of the skipped iterations if their number is odd.

The invariant the GPU tests check on real fits (test_fit_stops_exactly_where_the_reference_rule_fires, mp_worker.py
mmctm_fit) is checked here for EVERY batch size, stop position and both speculation outcomes: after the batch `lam` holds the
lambda of the stopping iteration j, `lam_prev` that of j - 1 and `sumtheta` the sum-theta of j (what the closing ELBO reads).
Without the second sum-theta buffer the speculative theta pass destroys the last of these (the bug the two-rank test found).
"""
import itertools

import pytest


class Host:
    def __init__(self, two_sumtheta_buffers=True):
        self.buf = {"A": 10, "B": 9, "S0": 10, "S1": 9}          # what each device buffer holds: the iteration that wrote it
        self.lam, self.lam_prev = "A", "B"                        # after iteration 10: lam = lambda_10, lam_prev = lambda_9
        self.st, self.st_alt = "S0", ("S1" if two_sumtheta_buffers else "S0")

    def run_batch(self, first, nb, stop_at, speculative_theta):
        """Iterations first .. first + nb - 1 are enqueued; the rule fires at the end of iteration stop_at (None: never)."""
        done, n_exec = False, 0
        for it in range(first, first + nb):
            self.lam, self.lam_prev = self.lam_prev, self.lam      # host-side swaps at enqueue time
            self.st, self.st_alt = self.st_alt, self.st
            theta_runs = (not done) or (speculative_theta and it == stop_at + 1)
            if theta_runs:
                self.buf[self.st] = it                             # the theta pass writes this iteration's sum-theta buffer
            if not done:
                assert self.buf[self.lam_prev] == it - 1           # the E-step reads lambda of the previous iteration
                self.buf[self.lam] = it                            # the solve writes lambda
                n_exec += 1
                if stop_at is not None and it == stop_at:
                    done = True
        if (nb - n_exec) & 1:                                      # the parity fix of mmctm_run_iterations
            self.lam, self.lam_prev = self.lam_prev, self.lam
            self.st, self.st_alt = self.st_alt, self.st
        return n_exec, done


@pytest.mark.parametrize("nb", range(1, 9))
def test_every_stop_position_leaves_the_state_of_the_stopping_iteration(nb):
    for stop_off, spec in itertools.product(list(range(nb)) + [None], (False, True)):
        h = Host()
        stop_at = None if stop_off is None else 11 + stop_off
        n_exec, done = h.run_batch(11, nb, stop_at, spec)
        j = stop_at if done else 10 + nb
        assert n_exec == j - 10
        assert h.buf[h.lam] == j and h.buf[h.lam_prev] == j - 1, (nb, stop_off, spec)
        assert h.buf[h.st] == j, (nb, stop_off, spec)              # the sum-theta the closing ELBO reads
        if not done:                                               # and a following batch starts from a consistent state
            n2, _ = h.run_batch(11 + nb, nb, None, spec)
            assert n2 == nb and h.buf[h.lam] == 10 + 2 * nb and h.buf[h.st] == 10 + 2 * nb


def test_one_sumtheta_buffer_is_not_enough_with_a_speculative_theta_pass():
    """The state before the second buffer existed: lambda is right, the sum-theta of the stopping iteration is gone."""
    h = Host(two_sumtheta_buffers=False)
    n_exec, done = h.run_batch(11, 8, 15, True)
    assert done and n_exec == 5 and h.buf[h.lam] == 15 and h.buf[h.lam_prev] == 14
    assert h.buf[h.st] == 16                                       # overwritten by the theta pass of iteration 16
    h = Host(two_sumtheta_buffers=False)
    h.run_batch(11, 8, 15, False)
    assert h.buf[h.st] == 15                                       # fine as long as nothing runs ahead of the decision
