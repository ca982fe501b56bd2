"""The library-wide constants of the reference (megastep/core.py:10-14), in a module of their own so that the CPU-only
scene-building code can use them without loading the native library."""
AGENT_WIDTH = .15
TEXTURE_RES = .05
# radius of the disc containing the agent: collision radius and near camera plane
AGENT_RADIUS = 1 / 2 ** .5 * AGENT_WIDTH


def gamma_encode(x):
    """linear -> viewable RGB (core.py:16-18)"""
    return x ** (1 / 2.2)


def gamma_decode(x):
    """viewable -> linear (interpolatable) RGB (core.py:20-22)"""
    return x ** 2.2
