#!/usr/bin/env Rscript
# reference_dump.R — run on ANY machine that has R >= 4.4 + multiDEGGs (the unmodified reference) + MASS,
# sfsmisc (+ qvalue): dumps per-edge t/F, p and p.adj of configs 1-2 (bundled data) so that TRUE-R goldens can
# replace the "numerically unpinned" status of oracle/ (SURVEY §8 c).  Not runnable in the build image (no R).
# Output: reference_goldens.tsv (one row per omic x method x category x edge).
suppressPackageStartupMessages(library(multiDEGGs))
data("synthetic_metadata"); data("synthetic_rnaseqData"); data("synthetic_proteomicData"); data("synthetic_OlinkData")
assays <- list(RNAseq = synthetic_rnaseqData, Proteomics = synthetic_proteomicData, Olink = synthetic_OlinkData)
rows <- list()
for (method in c("lm", "rlm")) for (padj in c("bonferroni", "BH")) {
  d <- get_diffNetworks(assays, synthetic_metadata, category_variable = "response", regression_method = method,
                        padj_method = padj, verbose = FALSE, show_progressBar = FALSE, cores = 1)
  for (omic in names(assays)) for (ctg in levels(d$metadata)) {
    net <- d$diffNetworks[[omic]][[ctg]]
    if (is.data.frame(net))
      rows[[length(rows) + 1]] <- data.frame(omic = omic, method = method, padj = padj, category = ctg,
                                             from = net$from, to = net$to,
                                             p.value = sprintf("%.17g", net$p.value),
                                             p.adj = sprintf("%.17g", net$p.adj))
  }
}
# single-pair statistics with the exact reference code path
g01 <- as.numeric(synthetic_metadata$response) - 1
for (e in list(c("FANCD2", "FAN1"), c("TNF", "TNFRSF1A"), c("TNF", "FAS"))) {
  p <- multiDEGGs:::fit_lm_interaction(synthetic_rnaseqData[e[1], ], synthetic_rnaseqData[e[2], ], g01)
  cat(sprintf("fit_lm_interaction %s->%s p = %.17g\n", e[1], e[2], p))
}
write.table(do.call(rbind, rows), "reference_goldens.tsv", sep = "\t", quote = FALSE, row.names = FALSE)
