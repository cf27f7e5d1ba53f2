"""Minimal `tf.contrib.framework.arg_scope` / `add_arg_scope` (used by models/gan.py:242-245,276-279):
an enclosing scope overrides keyword defaults of the decorated layer functions."""
import contextlib
import functools

_stack = [{}]


def add_arg_scope(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        merged = dict(_stack[-1].get(wrapper, {}))
        merged.update(kwargs)
        return fn(*args, **merged)

    wrapper._arg_scope_target = fn
    return wrapper


@contextlib.contextmanager
def arg_scope(fns, **kwargs):
    frame = {k: dict(v) for k, v in _stack[-1].items()}
    for f in fns:
        frame.setdefault(f, {}).update(kwargs)
    _stack.append(frame)
    try:
        yield
    finally:
        _stack.pop()
