"""Import shims that let an unmodified HermesPy run where its cluster / plotting / storage dependencies are absent.

    import hermespy_b200.shims as shims
    shims.install()      # before `import hermespy`; touches only packages that are really missing
"""
from .stubs import install  # noqa: F401
