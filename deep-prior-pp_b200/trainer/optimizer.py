"""Optimizer (reference: src/trainer/optimizer.py:39-117).  In the reference this class builds
Theano update expressions; here it is the descriptor the trainer hands to the engine, which runs
``dpp_adam_step`` over the flat parameter arena (same arithmetic: m, v, bias correction with
beta1**t / beta2**t, epsilon outside the sqrt, t starting at 1, fp32 constants)."""


class Optimizer(object):
    def __init__(self, grads, params):
        self.grads = grads
        self.params = params
        self.updates = []
        self.shared = []
        if grads is not None and len(grads) != len(params):
            print("Warning: Size of gradients ({}) does not fit size of parameters ({})!".format(len(grads), len(params)))

    def ADAM(self, learning_rate=0.0002, beta1=0.9, beta2=0.999, epsilon=1e-8, gamma=1 - 1e-8):
        if (beta1, beta2, epsilon) != (0.9, 0.999, 1e-8):
            raise NotImplementedError("dpp_adam_step implements the reference's default ADAM constants")
        self.updates = [('adam', dict(beta1=beta1, beta2=beta2, epsilon=epsilon, gamma=gamma))]
        return self.updates

    def RMSProp(self, learning_rate=0.01, decay=0.9, epsilon=1.0 / 100.):
        raise NotImplementedError("RMSProp is unused by the reference's entry scripts")
