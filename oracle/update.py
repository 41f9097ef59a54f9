"""Oracle (TEST INFRASTRUCTURE ONLY): rollout bookkeeping, n-step returns, clip, RMSProp, lr, sampling.

Restates:
  * ``ActorLearner.rescale_reward``        actor_learner.py:95-101
  * mask / bootstrap / n-step recurrence   paac.py:119, 140-149  (float64 NumPy in the reference)
  * ``tf.clip_by_global_norm`` as built at actor_learner.py:54-59 (op order from the shipped
    .meta): norm = sqrt(2 * sum_i L2Loss(g_i)), scale = clip * min(1/norm, 1/clip)
  * ``tf.train.RMSPropOptimizer`` -> ``ApplyRMSProp`` (actor_learner.py:33-34,70):
    ms += (g*g - ms) * (1 - rho); mom = momentum*mom + lr*g / sqrt(ms + eps); var -= mom
    (TF 1.0 training_ops functor; epsilon INSIDE the sqrt), slots ms=1, mom=0 (App. B)
  * ``ActorLearner.get_lr``                actor_learner.py:119-123
  * ``PAACLearner.__sample_policy_action`` paac.py:34-45, as inverse-CDF on injected uniforms
"""
import numpy as np


def rescale_reward(r):
    """actor_learner.py:95-101, vectorised."""
    return np.clip(np.asarray(r, np.float64), -1.0, 1.0)


def nstep_returns(rewards, episode_over, values, bootstrap_v, gamma):
    """paac.py:119,123,140-149.

    rewards[T,N] raw env rewards, episode_over[T,N] in {0,1}, values[T,N] f32 (acting V(s_t)),
    bootstrap_v[N] f32 (V(s_T)).  The reference holds all of these in float64 arrays and feeds
    the results to float32 placeholders, so: float64 recurrence, float32 outputs.
    Returns (y[T,N] f32, adv[T,N] f32).
    """
    T, N = rewards.shape
    r = rescale_reward(rewards)                                            # paac.py:123
    masks = 1.0 - np.asarray(episode_over, np.float32).astype(np.float64)  # paac.py:119
    vals = np.asarray(values, np.float32).astype(np.float64)               # paac.py:111 (f32 -> f64 array)
    R = np.asarray(bootstrap_v, np.float32).astype(np.float64).copy()      # paac.py:144
    y = np.zeros((T, N)); adv = np.zeros((T, N))
    for t in reversed(range(T)):                                           # paac.py:146-149
        R = r[t] + gamma * R * masks[t]
        y[t] = R
        adv[t] = R - vals[t]
    return y.astype(np.float32), adv.astype(np.float32)


def global_norm(grads):
    """sqrt(2 * sum_i (sum(g_i^2)/2)) in fp32 per tensor, as tf.global_norm does."""
    halves = [np.float32(np.sum(np.square(np.asarray(g, np.float32)), dtype=np.float32) / np.float32(2)) for g in grads]
    return np.float32(np.sqrt(np.float32(2.0) * np.sum(np.asarray(halves, np.float32), dtype=np.float32)))


def clip_by_global_norm(grads, clip):
    """actor_learner.py:54-59; returns (clipped list, norm).  fp32."""
    norm = global_norm(grads)
    with np.errstate(divide='ignore'):
        scale = np.float32(clip) * np.minimum(np.float32(1.0) / norm, np.float32(1.0) / np.float32(clip))
    return [np.asarray(g, np.float32) * scale for g in grads], norm


def rmsprop_apply(var, ms, mom, grad, lr, rho, eps, momentum=0.0):
    """TF-1.0 ApplyRMSProp, fp32, returns new (var, ms, mom)."""
    f = np.float32
    var, ms, mom, grad = (np.asarray(a, np.float32) for a in (var, ms, mom, grad))
    ms = ms + (grad * grad - ms) * (f(1.0) - f(rho))
    mom = mom * f(momentum) + (grad * f(lr)) / np.sqrt(ms + f(eps))
    var = var - mom
    return var.astype(np.float32), ms.astype(np.float32), mom.astype(np.float32)


def get_lr(global_step, initial_lr, lr_annealing_steps):
    """actor_learner.py:119-123 (python float arithmetic)."""
    if global_step <= lr_annealing_steps:
        return initial_lr - (global_step * initial_lr / lr_annealing_steps)
    return 0.0


def sample_actions(pi, uniforms):
    """Inverse-CDF categorical sampling on injected uniforms u in [0,1).

    Defines what "the same seeds" means for paac.py:34-45 (np.random.multinomial cannot be driven
    by injected uniforms): fp32 running sum c_j = c_{j-1} + pi_j in index order, action = first j
    with u < c_j, the last category is open-ended (absorbs rounding, as the reference's
    ``probs - epsneg`` trick makes multinomial's last bucket do).
    """
    pi = np.asarray(pi, np.float32)
    u = np.asarray(uniforms, np.float32)
    n, A = pi.shape
    act = np.full(n, A - 1, np.int32)
    c = np.zeros(n, np.float32)
    done = np.zeros(n, bool)
    for j in range(A - 1):
        c = (c + pi[:, j]).astype(np.float32)
        hit = (~done) & (u < c)
        act[hit] = j
        done |= hit
    return act
