#!/usr/bin/env Rscript
# run_reference.R <assay.f64> <G> <n> <genes.tsv> <group.tsv> [cores]
# Times the UNMODIFIED reference get_diffNetworks() on the same synthetic inputs bench.py uses (raw
# little-endian FP64 column-major matrix written by `python -m multideggs_b200.synth_dump`), with
# cores = parallel::detectCores() unless given.  Prints wall seconds and sum(tot tests).  Needs R + multiDEGGs;
# bench.py invokes it automatically iff `Rscript` is on PATH (it is not in the build image nor on the GPU box).
suppressPackageStartupMessages(library(multiDEGGs))
args <- commandArgs(trailingOnly = TRUE)
G <- as.integer(args[2]); n <- as.integer(args[3])
x <- matrix(readBin(args[1], "double", n = G * n, endian = "little"), nrow = G, ncol = n)
rownames(x) <- readLines(args[4])
grp <- read.delim(args[5], header = TRUE, stringsAsFactors = FALSE)
colnames(x) <- grp$sample
meta <- setNames(factor(grp$category), grp$sample)
cores <- if (length(args) >= 6) as.integer(args[6]) else parallel::detectCores()
t0 <- proc.time()[["elapsed"]]
d <- get_diffNetworks(x, meta, regression_method = "lm", padj_method = "BH", verbose = FALSE,
                      show_progressBar = FALSE, cores = cores)
dt <- proc.time()[["elapsed"]] - t0
cat(sprintf("{\"impl\": \"R\", \"seconds\": %.3f, \"cores\": %d, \"sig_pvalues_count\": %d}\n", dt, cores,
            as.integer(d$diffNetworks[[1]]$sig_pvalues_count)))
