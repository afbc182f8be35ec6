// employment.cuh — EmploymentEngine.apply_job_state (reference src/townlet/world/employment.py:63-142 and the helpers
// :229-455, :466-480) as a device step (included at the end of townlet_b200.cu; uses cuda_ok, set_error, g_launches).
//
// One THREAD per town walks the town's agents in order: the walk is sequential in the reference too — an agent's
// coworker look-up sees this tick's state for the agents before it and last tick's for the agents after it.  The step
// is a few dozen integer / float64 operations per agent on ~100 bytes of context, far off the hot path's cost; it is
// here so that the fields the observation and the reward read (wallet, wages, shift state, lateness, attendance,
// punctuality) can stay resident on the device across ticks.

namespace {

struct EmpAgent {  // the context of the agent being stepped, in registers
    int state, late_ticks, on_time_ticks, scheduled_ticks, shift_started_tick, shift_end_tick, last_present_tick;
    uint32_t cf;   // TL_EMP_F_*
    double wallet, wages_paid, wages_withheld;
    int lateness, late_ticks_today;
    bool on_shift;
    int shift_code;  // TL_F_SHIFT_* code of snapshot.shift_state
    uint32_t events;
};

__device__ __forceinline__ int emp_shift_code(int state) {  // snapshot.shift_state -> the feature's one-hot code (features.py:151-169)
    return state == TL_EMP_ON_TIME ? 1 : state == TL_EMP_LATE ? 2 : state == TL_EMP_ABSENT ? 3 : state == TL_EMP_POST_SHIFT ? 4 : 0;
}

// _employment_finalize_shift, employment.py:420-455
__device__ void emp_finalize(const TlTownBatch& b, const TlEmploymentParams& p, const TlEmploymentState& s, int a, EmpAgent& c) {
    if (!(c.cf & TL_EMP_F_SHIFT_STARTED) || (c.cf & TL_EMP_F_OUTCOME_RECORDED)) {
        c.shift_code = emp_shift_code(TL_EMP_POST_SHIFT);
        c.on_shift = false;
        return;
    }
    const int scheduled = max(1, c.scheduled_ticks);
    const double value = (double)c.on_time_ticks / (double)scheduled;
    double* samples = s.attendance_samples + (size_t)a * TL_MAX_ATTENDANCE_WINDOW;
    int n = s.attendance_count[a];
    const int window = min(max(1, p.attendance_window), TL_MAX_ATTENDANCE_WINDOW);
    if (n >= window) {  // deque(maxlen=window): the oldest sample drops out
        for (int i = 1; i < n; ++i) samples[i - 1] = samples[i];
        n = window - 1;
    }
    samples[n++] = value;
    s.attendance_count[a] = n;
    double total = 0.0;  // sum(...) adds left to right starting from 0
    for (int i = 0; i < n; ++i) total += samples[i];
    const_cast<double*>(b.attendance)[a] = total / (double)n;
    c.cf |= TL_EMP_F_OUTCOME_RECORDED;
    c.state = TL_EMP_POST_SHIFT;
    c.shift_code = emp_shift_code(TL_EMP_POST_SHIFT);
    c.on_shift = false;
    c.cf &= ~(TL_EMP_F_SHIFT_STARTED | TL_EMP_F_LAST_PRESENT | TL_EMP_F_LATE_PENALTY | TL_EMP_F_ABSENCE_PENALTY | TL_EMP_F_LATE_EVENT |
              TL_EMP_F_ABSENCE_EVENT | TL_EMP_F_DEPARTURE_EVENT | TL_EMP_F_LATE_COUNTER);
    c.late_ticks_today = c.late_ticks;
    c.late_ticks = 0;
}

// _employment_idle_state, employment.py:229-243
__device__ void emp_idle(const TlTownBatch& b, const TlEmploymentParams& p, const TlEmploymentState& s, int a, EmpAgent& c) {
    if (c.state != TL_EMP_PRE_SHIFT && c.state != TL_EMP_IDLE) emp_finalize(b, p, s, a, c);
    c.state = TL_EMP_IDLE;
    c.shift_code = 0;
    c.on_shift = false;
}

// _employment_coworkers_on_shift, employment.py:466-480: is any other agent of the town on the same job on_time or late?
__device__ bool emp_has_coworker(const TlEmploymentState& s, int a0, int a1, int a, int job) {
    bool any = false;
    for (int o = a0; o < a1; ++o) {
        if (o == a || s.job[o] != job) continue;
        const int st = s.state[o];
        any = any || st == TL_EMP_ON_TIME || st == TL_EMP_LATE;
    }
    return any;
}

__global__ void __launch_bounds__(128) employment_kernel(const TlTownBatch b, const TlEmploymentParams p, const TlEmploymentState s,
                                                         uint32_t* __restrict__ events_out, int n_ticks, int tick_offset) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= b.n_towns) return;
    const int a0 = b.town_agent_off[t], a1 = b.town_agent_off[t + 1];
    const int ox = s.town_origin ? s.town_origin[2 * t] : 0, oy = s.town_origin ? s.town_origin[2 * t + 1] : 0;
    const int tpd = max(1, p.ticks_per_day);
    const long long week = 7ll * tpd;
    double* wallet = const_cast<double*>(b.wallet);
    double* wages_withheld = const_cast<double*>(b.wages_withheld);
    double* wages_paid = const_cast<double*>(b.wages_paid);
    double* punctuality = const_cast<double*>(b.punctuality);
    int32_t* lateness = const_cast<int32_t*>(b.lateness);
#pragma unroll 1
    for (int it = 0; it < n_ticks; ++it) {
        const int tick = b.town_tick[t] + tick_offset + it;
        int day = tick / tpd;  // Python's floor division
        if (tick % tpd != 0 && tick < 0) --day;
#pragma unroll 1
        for (int a = a0; a < a1; ++a) {
            EmpAgent c;
            c.state = s.state[a], c.late_ticks = s.late_ticks[a], c.on_time_ticks = s.on_time_ticks[a];
            c.scheduled_ticks = s.scheduled_ticks[a], c.shift_started_tick = s.shift_started_tick[a];
            c.shift_end_tick = s.shift_end_tick[a], c.last_present_tick = s.last_present_tick[a], c.cf = s.ctx_flags[a];
            c.wallet = wallet[a], c.wages_paid = wages_paid[a], c.wages_withheld = wages_withheld[a];
            c.lateness = lateness[a], c.late_ticks_today = s.late_ticks_today[a];
            const uint32_t fl0 = b.flags[a];
            c.on_shift = (fl0 & TL_F_ON_SHIFT) != 0;
            c.shift_code = (int)((fl0 & TL_F_SHIFT_MASK) >> TL_F_SHIFT_SHIFT);
            c.events = 0;
            if (s.current_day[a] != day) {  // employment.py:73-78
                s.current_day[a] = day;
                c.late_ticks = 0;
                c.wages_paid = 0.0;
                c.late_ticks_today = 0;
            }
            int job = s.job[a];
            bool done = false;
            if (job < 0 || job >= p.n_jobs) {  // :80-86
                if (p.n_jobs <= 0) {
                    emp_idle(b, p, s, a, c);
                    done = true;
                } else {
                    job = 0;
                    s.job[a] = 0;
                }
            }
            if (!done) {
                const int start = p.job_start[job], end = p.job_end[job];
                const double wage_rate = p.job_wage[job], lateness_penalty = p.job_lateness_penalty[job];
                const int32_t pw = b.pos[a];
                const bool at_location = !p.job_has_location[job] || (pos_x(pw) + ox == p.job_x[job] && pos_y(pw) + oy == p.job_y[job]);
                {  // absence events older than seven days fall out (:103-107)
                    int32_t* ev = s.absence_events + (size_t)a * TL_MAX_ABSENCE_EVENTS;
                    int n = s.absence_count[a], drop = 0;
                    while (drop < n && (long long)tick - ev[drop] > week) ++drop;
                    if (drop) {
                        for (int i = drop; i < n; ++i) ev[i - drop] = ev[i];
                        n -= drop;
                        s.absence_count[a] = n;
                    }
                    s.absent_shifts_7d[a] = n;
                }
                if (tick < start - p.arrival_buffer) {
                    emp_idle(b, p, s, a, c);
                } else if (tick < start) {  // _employment_prepare_state, :245-248
                    c.state = TL_EMP_AWAIT_START;
                    c.shift_code = 0;
                    c.on_shift = false;
                } else if (tick <= end) {
                    // _employment_begin_shift, :250-268
                    if (!(c.cf & TL_EMP_F_SHIFT_STARTED) || c.shift_started_tick != start) {
                        c.shift_started_tick = start;
                        c.shift_end_tick = end;
                        c.scheduled_ticks = max(1, end - start + 1);
                        c.on_time_ticks = 0;
                        c.late_ticks = 0;
                        c.wages_paid = 0.0;
                        c.cf = TL_EMP_F_SHIFT_STARTED;  // every boolean key is reset, last_present_tick becomes None
                    }
                    // _employment_determine_state, :270-301
                    const int since = tick - start;
                    int state = c.state;
                    if (at_location) {
                        c.last_present_tick = tick;
                        c.cf |= TL_EMP_F_LAST_PRESENT;
                        if ((c.cf & TL_EMP_F_EVER_ON_TIME) || since <= p.grace_ticks) {
                            state = TL_EMP_ON_TIME;
                            c.cf |= TL_EMP_F_EVER_ON_TIME;
                        } else {
                            state = TL_EMP_LATE;
                        }
                    } else if (since <= p.grace_ticks) {
                        state = TL_EMP_LATE;
                    } else if (since > p.absent_cutoff) {
                        state = TL_EMP_ABSENT;
                    } else if ((c.cf & TL_EMP_F_LAST_PRESENT) && tick - c.last_present_tick > p.absence_slack) {
                        state = TL_EMP_ABSENT;
                    } else {
                        state = TL_EMP_LATE;
                    }
                    // _employment_apply_state_effects, :303-418
                    const int previous = c.state;
                    c.state = state;
                    s.state[a] = state;  // (the coworker look-ups of later agents read it)
                    c.shift_code = emp_shift_code(state);
                    const bool eligible = (state == TL_EMP_ON_TIME || state == TL_EMP_LATE) && at_location;
                    const bool coworkers = emp_has_coworker(s, a0, a1, a, job);
                    const bool previous_working = previous == TL_EMP_ON_TIME || previous == TL_EMP_LATE;
                    if (state == TL_EMP_ON_TIME) c.on_time_ticks += 1;
                    if (state == TL_EMP_LATE) {
                        c.late_ticks += 1;
                        c.late_ticks_today += 1;
                        if (!(c.cf & TL_EMP_F_LATE_PENALTY) && !previous_working) {
                            c.wallet = fmax(0.0, c.wallet - lateness_penalty);
                            c.cf |= TL_EMP_F_LATE_PENALTY;
                            if (!(c.cf & TL_EMP_F_LATE_COUNTER)) {
                                c.lateness += 1;
                                c.cf |= TL_EMP_F_LATE_COUNTER;
                            }
                        }
                        if (!(c.cf & TL_EMP_F_LATE_EVENT)) {
                            c.events |= TL_EMP_EV_LATE_START;
                            c.cf |= TL_EMP_F_LATE_EVENT;
                        }
                        if (!at_location) {
                            if (p.late_tick_penalty != 0.0) c.wallet = fmax(0.0, c.wallet - p.late_tick_penalty);
                            c.wages_withheld += wage_rate;
                        }
                        if (coworkers && !(c.cf & TL_EMP_F_LATE_HELP_EVENT)) {
                            c.events |= TL_EMP_EV_HELPED_WHEN_LATE;  // the host applies update_relationship(..., "employment_help")
                            c.cf |= TL_EMP_F_LATE_HELP_EVENT;
                        }
                    }
                    if (state == TL_EMP_ABSENT) {
                        if (!(c.cf & TL_EMP_F_ABSENCE_PENALTY)) {
                            c.wallet = fmax(0.0, c.wallet - p.absence_penalty);
                            c.cf |= TL_EMP_F_ABSENCE_PENALTY;
                        }
                        c.wages_withheld += wage_rate;
                        if (!(c.cf & TL_EMP_F_ABSENCE_EVENT)) {
                            c.events |= TL_EMP_EV_ABSENT;
                            c.cf |= TL_EMP_F_ABSENCE_EVENT;
                            int32_t* ev = s.absence_events + (size_t)a * TL_MAX_ABSENCE_EVENTS;
                            int n = s.absence_count[a];
                            if (n >= TL_MAX_ABSENCE_EVENTS) {  // (more absences inside one week than the table holds: keep the newest)
                                for (int i = 1; i < n; ++i) ev[i - 1] = ev[i];
                                n = TL_MAX_ABSENCE_EVENTS - 1;
                            }
                            ev[n++] = tick;
                            s.absence_count[a] = n;
                            s.absent_shifts_7d[a] = n;
                        }
                        if (coworkers && !(c.cf & TL_EMP_F_TOOK_SHIFT_EVENT)) {
                            c.events |= TL_EMP_EV_TOOK_MY_SHIFT;  // update_relationship(..., "employment_shift_taken") on the host
                            c.cf |= TL_EMP_F_TOOK_SHIFT_EVENT;
                        }
                    } else if (state == TL_EMP_ON_TIME || state == TL_EMP_LATE) {
                        c.cf &= ~TL_EMP_F_ABSENCE_PENALTY;
                    }
                    if (eligible && state != TL_EMP_ABSENT) {
                        c.on_shift = true;
                        c.wallet += wage_rate;
                        s.wages_earned[a] += 1;
                        c.wages_paid += wage_rate;
                    } else {
                        c.on_shift = false;
                        if (previous_working && state != TL_EMP_ON_TIME && state != TL_EMP_LATE) {
                            if (!(c.cf & TL_EMP_F_DEPARTURE_EVENT) && previous != TL_EMP_ABSENT) {
                                c.events |= TL_EMP_EV_DEPARTED_EARLY;
                                c.cf |= TL_EMP_F_DEPARTURE_EVENT;
                            }
                        }
                    }
                } else {
                    emp_finalize(b, p, s, a, c);
                }
            }
            // write the context and the snapshot fields back
            s.state[a] = c.state, s.late_ticks[a] = c.late_ticks, s.on_time_ticks[a] = c.on_time_ticks;
            s.scheduled_ticks[a] = c.scheduled_ticks, s.shift_started_tick[a] = c.shift_started_tick;
            s.shift_end_tick[a] = c.shift_end_tick, s.last_present_tick[a] = c.last_present_tick, s.ctx_flags[a] = c.cf;
            s.late_ticks_today[a] = c.late_ticks_today;
            wallet[a] = c.wallet, wages_paid[a] = c.wages_paid, wages_withheld[a] = c.wages_withheld, lateness[a] = c.lateness;
            b.flags[a] = (b.flags[a] & ~(TL_F_ON_SHIFT | TL_F_SHIFT_MASK)) | (c.on_shift ? TL_F_ON_SHIFT : 0u) |
                         ((uint32_t)c.shift_code << TL_F_SHIFT_SHIFT);
            {  // employment_context_punctuality, employment.py:215-222
                const double value = (double)c.on_time_ticks / (double)max(1, c.scheduled_ticks);
                punctuality[a] = fmax(0.0, fmin(1.0, value));
            }
            if (events_out) events_out[(size_t)it * b.n_agents + a] = c.events;
        }
    }
}

}  // namespace

extern "C" int tl_employment_step(const TlTownBatch* batch, const TlEmploymentParams* params, const TlEmploymentState* state,
                                  uint32_t* events_out, int32_t n_ticks, int32_t tick_offset, tl_stream_t stream) {
    if (!batch || !params || !state) return set_error("tl_employment_step: NULL pointer"), TL_ERR_ARG;
    if (n_ticks < 0 || batch->n_towns < 0 || batch->n_agents < 0) return set_error("tl_employment_step: negative size"), TL_ERR_ARG;
    if (params->n_jobs < 0 || params->n_jobs > TL_MAX_JOBS) return set_error("tl_employment_step: more than TL_MAX_JOBS jobs"), TL_ERR_UNSUPPORTED;
    if (params->attendance_window > TL_MAX_ATTENDANCE_WINDOW) return set_error("tl_employment_step: attendance window above 14"), TL_ERR_UNSUPPORTED;
    if (batch->n_towns == 0 || batch->n_agents == 0 || n_ticks == 0) return TL_OK;
    if (!state->job || !state->state || !state->current_day || !state->late_ticks || !state->ctx_flags || !state->shift_started_tick ||
        !state->shift_end_tick || !state->last_present_tick || !state->scheduled_ticks || !state->on_time_ticks ||
        !state->attendance_samples || !state->attendance_count || !state->absence_events || !state->absence_count ||
        !state->late_ticks_today || !state->absent_shifts_7d || !state->wages_earned)
        return set_error("tl_employment_step: a context array is NULL"), TL_ERR_ARG;
    if (!batch->wallet || !batch->wages_withheld || !batch->wages_paid || !batch->punctuality || !batch->lateness || !batch->attendance ||
        !batch->flags || !batch->pos || !batch->town_tick || !batch->town_agent_off)
        return set_error("tl_employment_step: a batch array is NULL"), TL_ERR_ARG;
    const int threads = 128;
    employment_kernel<<<(batch->n_towns + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(*batch, *params, *state, events_out,
                                                                                                      n_ticks, tick_offset);
    if (!cuda_ok(cudaGetLastError(), "employment_kernel launch")) return TL_ERR_CUDA;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return TL_OK;
}
