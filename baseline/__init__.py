"""Baseline arm: loader + offline install recipe for the unmodified reference (baseline/_ref is git-ignored)."""
