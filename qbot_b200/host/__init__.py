"""Host-side mirror of the reference's operator interface for the state path.

The reference's interpreter, ProbVal class and expression namespace are unchanged when this
backend is installed into a real qbot checkout (``qbot_b200.install()``).  Where the reference
is not importable (the GPU box), this sub-package supplies the same interface -- same op
names, argument meaning, error text and ProbVal semantics -- so that DSL programs and the
reference's own test programs run against the CUDA backend.
"""
