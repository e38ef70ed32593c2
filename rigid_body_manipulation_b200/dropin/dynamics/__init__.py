# same re-export shape as the reference's dynamics/__init__.py:1
from .dynamics import *  # noqa: F401,F403
