"""Drop-in replacement for the reference package `lib` (criterion + mixup entry points)."""
