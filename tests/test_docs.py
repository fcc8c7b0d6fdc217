"""The Python blocks of docs/MIGRATING.md are executable documentation."""
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _blocks():
    text = open(os.path.join(REPO, "docs", "MIGRATING.md")).read()
    return re.findall(r"```python\n(.*?)```", text, flags=re.S)


@pytest.mark.parametrize("idx", range(len(_blocks())))
def test_migrating_guide_block_runs(idx):
    exec(compile(_blocks()[idx], f"docs/MIGRATING.md[block {idx}]", "exec"), {"__name__": "__docs__"})
