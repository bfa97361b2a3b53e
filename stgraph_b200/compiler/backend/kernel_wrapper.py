"""Framework-neutral half of the autograd bridge (``stgraph/compiler/backend/kernel_wrapper.py:1-19``).

``forward(executor, kid, kernel_args, rets, *tensors)`` -> ``executor.forward_cb``;
``setup_context`` stores ``(executor, kid)`` and disables gradient materialisation;
``backward`` returns ``(None,)*4`` for the four non-tensor arguments followed by one gradient per
tensor argument.
"""


class KernelWrapper:
    @staticmethod
    def forward(executor, kid, kernel_args, rets, *args):
        return executor.forward_cb(kid, kernel_args, rets, args)

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.backward_cache = inputs[0], inputs[1]
        ctx.set_materialize_grads(False)

    @staticmethod
    def backward(ctx, *gradout):
        executor, kid = ctx.backward_cache
        return (None, None, None, None) + executor.backward_cb(kid, gradout)
