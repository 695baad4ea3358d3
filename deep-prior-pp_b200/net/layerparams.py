"""LayerParams base (reference: src/net/layerparams.py:35-104): input / output dimensions that re-derive dependent
shapes when reassigned, a printable activation name, the activation's output range.

The reference writes one property + setter pair per attribute; here ``tracked`` generates them."""
import inspect
import numpy


def tracked(attr, refresh=True, convert=None):
    """Property stored in ``_<attr>``.  Assigning it re-derives the dependent dimensions through ``update()`` (the
    entry scripts poke e.g. ``cfgParams.outputDim`` after construction, main_nyu_posereg_embedding.py:160-162)."""
    slot = '_' + attr

    def fget(self):
        return getattr(self, slot)

    def fset(self, value):
        setattr(self, slot, convert(value) if convert is not None else value)
        if refresh:
            self.update()
    return property(fget, fset)


_RANGES = {'tanh': [-1, 1], 'sigmoid': [0, 1], 'ReLU': [0, numpy.inf]}


class LayerParams(object):
    inputDim = tracked('inputDim')
    outputDim = tracked('outputDim')

    def __init__(self, inputDim, outputDim):
        self._inputDim, self._outputDim = inputDim, outputDim

    def update(self):
        """derived parameter classes recompute their shapes here"""

    @property
    def activation_str(self):
        """name the layer descriptions and the initialisation lookup use (layerparams.py:69-83)"""
        if not hasattr(self, 'activation'):
            return ''
        act = self.activation
        if act is None:
            return 'None'
        if inspect.isclass(act):
            return act.__class__.__name__          # (sic) the reference names the metaclass for classes
        return act.__name__ if inspect.isfunction(act) else str(act)

    def getOutputRange(self):
        unbounded = [-numpy.inf, numpy.inf]
        if not hasattr(self, 'activation'):
            return unbounded
        return list(_RANGES.get(self.activation_str, unbounded))
