"""Agent interface (reference: mbt_gym/agents/Agent.py:6-9): map an observation matrix to an action matrix."""


class Agent:
    def get_action(self, state):
        raise NotImplementedError

    def to_policy(self, env):
        """Optional: the agent as an `mbt_policy` evaluated on the device by the fused rollout kernel."""
        raise NotImplementedError(f"{type(self).__name__} has no on-device form")
