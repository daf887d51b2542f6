"""Bindings of the B200 path onto other code bases (import-guarded: nothing heavy is imported until install() is called)."""
