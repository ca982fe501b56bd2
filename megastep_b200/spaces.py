"""Shape-only descriptors of observation and action spaces (reference: megastep/spaces.py:3-28)."""


class _Space:
    shape = ()

    def __repr__(self):
        return f'{type(self).__name__}{self.shape}'


class MultiEmpty(_Space):
    pass


class MultiVector(_Space):
    """`dim` floats per agent."""

    def __init__(self, n_agents, dim):
        self.shape = (n_agents, dim)


class MultiImage(_Space):
    """A (C, H, W) image per agent."""

    def __init__(self, n_agents, C, H, W):
        self.shape = (n_agents, C, H, W)


class MultiConstant(_Space):
    def __init__(self, n_agents):
        self.shape = (n_agents,)


class MultiDiscrete(_Space):
    """One of `n_actions` choices per agent."""

    def __init__(self, n_agents, n_actions):
        self.shape = (n_agents, n_actions)
