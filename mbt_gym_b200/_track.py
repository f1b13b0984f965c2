"""Edit tracking for the host-side model objects.

The reference reads its Python attributes at every step, so `env.max_inventory = 5` or
`env.reward_function.per_step_inventory_aversion = 0.1` take effect immediately.  Here the parameters live in a device
handle that is (re)configured from the Python attributes; re-flattening them on every `step()` would cost more than the
ctypes call itself.  Instead every public attribute assignment on an environment, a model-dynamics object, a stochastic
process or a reward function bumps ONE process-wide version number; `TradingEnvironment.step()` compares it with the
version it last validated against (one integer comparison) and re-flattens only when something was edited.
"""
version = [0]


class Tracked:
    """Mixin: public attribute assignments bump the process-wide edit version (private `_names` do not)."""

    def __setattr__(self, name, value):
        if name[:1] != "_":
            version[0] += 1
        object.__setattr__(self, name, value)
