"""package marker (see ../../README.md)"""
