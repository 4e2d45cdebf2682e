"""The reference's obstacle-set fixtures (inputs only): test/obstaclesets/2D.jl:3-42 and
test/obstaclesets/ND.jl:1-14, as host-side shape objects / box lists."""
import numpy as np

from .shapes2d import Box2D, Circle, Compound2D, Polygon


def ISRR_2H():
    """test/obstaclesets/2D.jl:3-11"""
    return Compound2D([
        Box2D([.0, .16], [.36, .5]),
        Box2D([.4, .5], [.19, .35]),
        Box2D([.22, .46], [.57, .75]),
        Box2D([.75, 1.], [.64, .77]),
        Box2D([.22, .8], [.34, .51]),
    ])


def TRI_BALLS():
    """test/obstaclesets/2D.jl:13-20"""
    return Compound2D([
        Polygon([(.3, .3), (.7, .3), (.5, .65)]),
        Circle([.3, .3], .15),
        Circle([.7, .3], .15),
        Circle([.5, .65], .15),
    ])


def ISRR_POLY():
    """test/obstaclesets/2D.jl:22-30"""
    return Compound2D([
        Polygon([(.0, .25), (.27, .28), (.17, .4), (.0, .4)]),
        Polygon([(.5, .2), (.2, .5), (.25, .7), (.4, .8), (.6, .8), (.7, .5)]),
        Polygon([(.55, .2), (.75, .5), (.85, .5), (.85, .2)]),
        Circle([.9, .65], .1),
    ])


def ISRR_POLY_WITH_SPIKE():
    """test/obstaclesets/2D.jl:32-40"""
    return Compound2D([
        Polygon([(.0, .25), (.27, .28), (.17, .4), (.0, .4)]),
        Polygon([(.5, .2), (.2, .5), (.25, .7), (.4, .8), (.6, .8), (.7, .5)]),
        Polygon([(.55, .2), (.75, .5), (.85, .5), (.85, .2)]),
        Polygon([(.3, .6), (.15, .85), (.4, .6)]),
        Circle([.9, .65], .1),
    ])


def EMPTY_2D():
    """test/obstaclesets/2D.jl:42"""
    return Compound2D([])


def BOXES2D():
    """test/obstaclesets/ND.jl:1 -- each entry is a d x 2 [lo hi] matrix"""
    return [np.array(b, dtype=np.float64) for b in (
        [[0., 0.16], [0.36, 0.5]], [[0.4, 0.5], [0.19, 0.35]], [[0.22, 0.46], [0.57, 0.75]],
        [[0.75, 1.], [0.64, 0.77]], [[0.22, 0.8], [0.34, 0.51]])]


def BOXES3D():
    """test/obstaclesets/ND.jl:3-14 (rows lo;hi, transposed to d x 2)"""
    rows = [
        [[0.25, 0, 0], [.3, .4, 1]],
        [[0.25, .6, 0], [.3, 1, 1]],
        [[0.25, .4, 0], [.3, .6, .25]],
        [[0.25, .4, .33], [.3, .6, .7]],
        [[0.25, .4, .85], [.3, .6, 1]],
        [[0.7, 0, 0], [.75, 1, .3]],
        [[0.7, 0, .5], [.75, 1, 1]],
        [[0.7, 0, .3], [.75, .2, .5]],
        [[0.7, .4, .3], [.75, .5, .5]],
        [[0.7, .7, .3], [.75, 1, .5]],
    ]
    return [np.array(r, dtype=np.float64).T.copy() for r in rows]
