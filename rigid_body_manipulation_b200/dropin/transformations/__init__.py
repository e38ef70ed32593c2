# same re-export shape as the reference's transformations/__init__.py:1-2
from .poses import *  # noqa: F401,F403
from .transformations import *  # noqa: F401,F403
