# gpu_backend.R — drop-in body for multiDEGGs:::get_diffNetworks_singleOmic (reference
# R/core_functions.R:258-529) that keeps the R API and the `deggs` layout unchanged and replaces the
# per-percentile mclapply loop (core_functions.R:372-470) plus calc_pvalues_percentile /
# calc_pvalues_network / fit_lm_interaction (core_functions.R:702-1074) by ONE `.Call` into the CUDA
# engine (src/rcall.c -> libmdeggs, include/mdeggs.h).
#
# What stays in R (strings, factors, messages, data.frames): sample alignment, edge selection and
# de-duplication, `levels(metadata)`, rownames `from-to`, and every
# message()/stop() text.  Every `padj_method` is adjusted on the device; only hommel on more than 65,536 edges, and
# q.value when options(multiDEGGs.qvalue_on_host = TRUE), are run by R's own functions on the raw p-values of each
# (percentile, category) segment.  `cores` and `show_progressBar` are accepted and ignored by the GPU path except that
# `cores == 1` keeps the reference's error semantics for failing pairs (a failing pair aborts the omic).
#
# Install: see INTEGRATION.md.  NAMESPACE needs `useDynLib(multiDEGGs, .registration = TRUE)`.

.mdeggs_padj_code <- function(padj_method, n_edges = 0L) {
  if (padj_method == "none") return(0L)
  if (padj_method == "bonferroni") return(1L)
  if (padj_method %in% c("BH", "fdr")) return(2L)
  if (padj_method == "holm") return(4L)
  if (padj_method == "hochberg") return(5L)
  if (padj_method == "BY") return(6L)
  # q.value: the device estimates pi0 (smoother over lambda = seq(.05,.95,.05)) and returns q-values (code 7).
  # options(multiDEGGs.qvalue_on_host = TRUE) keeps Bioconductor's qvalue::qvalue in the loop instead (code 3).
  if (padj_method == "q.value" && !isTRUE(getOption("multiDEGGs.qvalue_on_host"))) return(7L)
  # hommel: on the device up to 65,536 edges (MDEGGS_HOMMEL_MAX_E; it is O(n^2) per segment), else by R (code 3)
  if (padj_method == "hommel" && n_edges <= 65536L) return(8L)
  3L  # big hommel jobs or host-side q.value: raw p + masks come back, R adjusts each segment
}

# Host preparation of one layer (core_functions.R:280-340): everything up to the engine call.  Returns the argument
# list of the `.Call` (first eleven entries of mdeggs_R_run without n_gpus, in mdeggs_R_run_layers order) plus what
# the assembly needs afterwards.
.mdeggs_prepare_layer <- function(assayData, assayDataName, metadata, regression_method, mixedModel, network,
                                  percentile_vector, padj_method, verbose) {
  if (verbose) message("Generating ", assayDataName, " network layer...")
  if (mixedModel) stop("mixedModel = TRUE is not supported by the GPU engine; use the CPU implementation")
  if (!(is.matrix(assayData) || is.data.frame(assayData)))
    stop(assayDataName, " is not a matrix or data.frame")
  if (!any(colnames(assayData) %in% names(metadata)))
    stop("Sample IDs in ", assayDataName, " don't match the IDs in metadata
          (the metadata names/rownames must match the sample IDs used as column
          names of ", assayDataName, ".")
  assayData <- as.matrix(assayData)
  storage.mode(assayData) <- "double"
  assayData <- assayData[, which(colnames(assayData) %in% names(metadata)), drop = FALSE]
  assayData <- assayData[, colSums(is.na(assayData)) < (nrow(assayData) - 1), drop = FALSE]
  missing_ids <- names(metadata)[!(names(metadata) %in% colnames(assayData))]
  if (length(missing_ids) != 0)
    message("The following samples IDs are missing in ", assayDataName, ":\n",
            paste(missing_ids, collapse = ", "))
  metadata <- metadata[colnames(assayData)]
  if (length(unique(metadata)) == 1)
    stop("All sample IDs in ", assayDataName, " belong to one category.
           No differential analysis is possible.")

  network_to_use <- if (!is.null(network)) network else omic_network
  from_idx <- match(network_to_use$from, rownames(assayData))
  to_idx <- match(network_to_use$to, rownames(assayData))
  keep <- !is.na(from_idx) & !is.na(to_idx)
  if (!any(keep))
    stop("Rownames of ", assayDataName, " don't match any link of the biological
    network. Check if names need conversion to gene symbols or entrez IDs.
    Alternatively, provide more data.")
  from_idx <- from_idx[keep]; to_idx <- to_idx[keep]
  dup <- duplicated(cbind(from_idx, to_idx))            # paste(from, to) keep-first, on integer ids
  from_idx <- from_idx[!dup]; to_idx <- to_idx[!dup]

  categories <- levels(metadata)
  K <- length(categories)
  padj_code <- .mdeggs_padj_code(padj_method, length(from_idx))
  list(args = list(assayData, as.integer(from_idx), as.integer(to_idx),
                   as.integer(metadata), as.integer(K), as.double(percentile_vector),
                   as.integer(regression_method == "rlm"), padj_code,
                   3L,                                    # f.robftest(var = 3), core_functions.R:920
                   as.integer(getOption("multiDEGGs.rlm_cov", 0L)),  # 0 = XtWX, 1 = XtX
                   FALSE),                                # want_all
       gene_names = rownames(assayData), from_idx = from_idx, to_idx = to_idx, categories = categories,
       padj_code = padj_code)
}

# Engine results of one layer -> the reference's per-omic list (core_functions.R:471-529)
.mdeggs_assemble_layer <- function(prep, res, percentile_vector, padj_method, verbose, cores) {
  from_idx <- prep$from_idx; to_idx <- prep$to_idx
  categories <- prep$categories; K <- length(categories)
  padj_code <- prep$padj_code
  P <- length(percentile_vector)
  E <- length(from_idx)
  sig <- res$sig_count
  if (padj_code == 3L) {                                  # host-side adjustment per (percentile, category)
    cat_all <- matrix(as.integer(res$cat), nrow = E, ncol = P)
    padj_all <- matrix(NA_real_, nrow = E, ncol = P)
    sig <- rep(NA_real_, P)
    for (p in which(res$tot_edges > 0)) {
      n_sig <- 0
      for (k in seq_len(K)) {
        rows <- which(cat_all[, p] == k)
        if (length(rows) == 0) next
        pv <- res$p_edge[rows]
        if (padj_method == "q.value") {
          adj <- try(qvalue::qvalue(pv)$qvalues, silent = TRUE)
          if (methods::is(adj, "try-error"))
            adj <- if (length(rows) > 1) qvalue::qvalue(p = pv, pi0 = 1)$qvalues else NA
        } else {
          adj <- try(stats::p.adjust(pv, method = padj_method))
          if (methods::is(adj, "try-error")) adj <- NA
        }
        padj_all[rows, p] <- adj
        n_sig <- n_sig + sum(adj < 0.05, na.rm = TRUE)
      }
      sig[p] <- n_sig
    }
  }
  valid <- which(res$tot_edges > 0)
  failed <- valid[res$fail_count[valid] > 0]
  if (length(failed) > 0) {
    if (cores == 1) stop("system is computationally singular or NA/NaN/Inf in a tested gene pair")
    valid <- setdiff(valid, failed)                       # mclapply: only that percentile is lost
  }
  if (length(valid) == 0) stop("No significant differences detected across categories.")
  best <- valid[which.max(sig[valid])]                    # first maximum
  if (verbose)
    message("Percolation analysis: genes whose expression is below the ",
            percentile_vector[best] * 100, "th percentile are removed from networks.")
  if (padj_code == 3L) {
    cat_best <- cat_all[, best]; padj_best <- padj_all[, best]
  } else {
    cat_best <- as.integer(res$cat_best); padj_best <- res$padj_best
  }
  gene_names <- prep$gene_names
  final_networks <- lapply(seq_len(K), function(k) {
    rows <- which(cat_best == k)
    if (length(rows) == 0) return("No specific links for this category.")
    df <- data.frame(from = gene_names[from_idx[rows]], to = gene_names[to_idx[rows]],
                     p.value = res$p_edge[rows], stringsAsFactors = FALSE)
    if (padj_method != "none") df$p.adj <- padj_best[rows]
    rownames(df) <- paste(df$from, df$to, sep = "-")
    df
  })
  names(final_networks) <- categories
  final_networks <- append(final_networks, sig[best])
  names(final_networks)[length(final_networks)] <- "sig_pvalues_count"
  final_networks
}

get_diffNetworks_singleOmic <- function(assayData, assayDataName, metadata, regression_method,
                                        mixedModel, id_variable, lmer_ctrl, network,
                                        percentile_vector, padj_method, show_progressBar,
                                        verbose, cores) {
  prep <- .mdeggs_prepare_layer(assayData, assayDataName, metadata, regression_method, mixedModel, network,
                                percentile_vector, padj_method, verbose)
  a <- prep$args
  res <- .Call("mdeggs_R_run", a[[1]], a[[2]], a[[3]], a[[4]], a[[5]], a[[6]], a[[7]], a[[8]], a[[9]], a[[10]],
               as.integer(getOption("multiDEGGs.gpus", 1L)), a[[11]])
  .mdeggs_assemble_layer(prep, res, percentile_vector, padj_method, verbose, cores)
}

# Drop-in for the per-omic lapply of get_diffNetworks (core_functions.R:214-234): every layer is prepared on the
# host, ALL layers run as one launch set (one `.Call`, mdeggs_run_omics: one engine context per layer, the layers
# overlap on the GPU), then each layer is assembled.  A layer that fails in any of the three steps is reported with
# message() and becomes NULL, exactly like the reference's per-omic tryCatch (core_functions.R:219-233).
.mdeggs_diffNetworks_layers <- function(assayDataList, metadata, regression_method, mixedModel, network,
                                        percentile_vector, padj_method, verbose, cores) {
  preps <- lapply(names(assayDataList), function(nm)
    tryCatch(.mdeggs_prepare_layer(assayDataList[[nm]], nm, metadata, regression_method, mixedModel, network,
                                   percentile_vector, padj_method, verbose),
             error = function(e) { message(conditionMessage(e)); NULL }))
  ok <- which(!vapply(preps, is.null, logical(1)))
  res <- if (length(ok)) .Call("mdeggs_R_run_layers", lapply(preps[ok], `[[`, "args")) else list()
  out <- vector("list", length(assayDataList))
  names(out) <- names(assayDataList)
  for (j in seq_along(ok)) {
    i <- ok[j]
    val <- tryCatch({
      if (is.character(res[[j]])) stop(res[[j]])        # engine error of this layer
      .mdeggs_assemble_layer(preps[[i]], res[[j]], percentile_vector, padj_method, verbose, cores)
    }, error = function(e) { message(conditionMessage(e)); NULL })
    if (!is.null(val)) out[[i]] <- val
  }
  out
}

# feature construction of predict.multiDEGGs_filter (feature_selection_for_ML.R:423-431) on the GPU
.mdeggs_pair_features <- function(newdata, pairs, ratio = TRUE) {
  x <- as.matrix(newdata); storage.mode(x) <- "double"
  a <- match(pairs[, 1], colnames(x)) - 1L
  b <- match(pairs[, 2], colnames(x)) - 1L
  out <- .Call("mdeggs_R_pair_features", x, as.integer(a), as.integer(b), ratio)
  colnames(out) <- paste(pairs[, 1], pairs[, 2], sep = ":")
  rownames(out) <- rownames(x)
  out
}

# per-category regression lines of plot_regressions (plotting_functions.R:125-172) from the moments the last engine run
# left on the device: list(intercept, slope, sigma, n, mean_x, sxx) per category, enough for abline() and for the
# confidence band  fit +- qt(0.975, n - 2) * sigma * sqrt(1 / n + (x0 - mean_x)^2 / sxx)
.mdeggs_pair_lines <- function(from_idx, to_idx, categories) {
  K <- length(categories)
  m <- .Call("mdeggs_R_pair_moments", as.integer(from_idx), as.integer(to_idx), as.integer(K))
  lapply(seq_along(from_idx), function(j) {
    out <- lapply(seq_len(K), function(k) {
      v <- m[, k, j]
      slope <- v[6] / v[4]
      list(intercept = v[3] - slope * v[2], slope = slope, n = v[1], mean_x = v[2], sxx = v[4],
           sigma = sqrt((v[5] - slope * v[6]) / (v[1] - 2)))
    })
    names(out) <- categories
    out
  })
}
