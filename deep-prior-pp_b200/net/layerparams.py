"""LayerParams base (reference: src/net/layerparams.py:35-104)."""
import inspect
import numpy


class LayerParams(object):
    def __init__(self, inputDim, outputDim):
        self._inputDim = inputDim
        self._outputDim = outputDim

    @property
    def outputDim(self):
        return self._outputDim

    @outputDim.setter
    def outputDim(self, value):
        self._outputDim = value
        self.update()

    @property
    def inputDim(self):
        return self._inputDim

    @inputDim.setter
    def inputDim(self, value):
        self._inputDim = value
        self.update()

    def update(self):
        pass

    @property
    def activation_str(self):
        # layerparams.py:69-83
        if hasattr(self, 'activation'):
            if self.activation is None:
                return str(None)
            elif inspect.isclass(self.activation):
                return self.activation.__class__.__name__
            elif inspect.isfunction(self.activation):
                return self.activation.__name__
            else:
                return str(self.activation)
        return ''

    def getOutputRange(self):
        if not hasattr(self, 'activation'):
            return [-numpy.inf, numpy.inf]
        s = self.activation_str
        if s == 'tanh':
            return [-1, 1]
        if s == 'sigmoid':
            return [0, 1]
        if s == 'ReLU':
            return [0, numpy.inf]
        return [-numpy.inf, numpy.inf]
