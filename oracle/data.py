"""Seeded synthetic inputs for the checker: the generator itself is product code (`main.py --synthetic`,
`bench.py`), so it lives in `continual_learning_b200/synthetic.py`; tests may keep importing it from here."""
from continual_learning_b200.synthetic import structured_batch, uniform_batch  # noqa: F401
