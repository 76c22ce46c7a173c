"""Drop-in `dataset` package.  The fork's .gitignore drops its whole dataset/ directory except gta5_dataset.py
(SURVEY.md Q3): cityscapes_dataset.py and the list files do not exist in the reference checkout.  The two classes here
keep the constructor keywords and item tuples the scripts use and serve the real files when their list file exists,
synthetic tensors of the same shapes and dtypes otherwise (no datasets are available offline)."""
