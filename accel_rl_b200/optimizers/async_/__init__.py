"""The reference package is named `async`, a Python keyword since 3.7; the
importable name here is `async_`, exporting the reference's class names."""
