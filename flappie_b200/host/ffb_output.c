/* ffb_output.c -- fasta / fastq / SAM records with the reference's exact text (src/flappie_output.c:92-133). */
#include <string.h>

#include "ffb_host.h"

static const char *const FORMAT_NAMES[] = {"fasta", "fastq", "sam"};

enum ffb_outformat ffb_get_outformat(const char *name) {
    if (name)
        for (int f = 0; f < (int)FFB_OUT_INVALID; f++)
            if (0 == strcmp(name, FORMAT_NAMES[f])) return (enum ffb_outformat)f;
    return FFB_OUT_INVALID;
}

const char *ffb_outformat_string(enum ffb_outformat fmt) {
    return ((int)fmt >= 0 && fmt < FFB_OUT_INVALID) ? FORMAT_NAMES[fmt] : NULL;
}

/* the JSON-ish comment both fasta and fastq headers carry */
static void header_line(FILE *fp, char lead, const char *name, const char *readname, const char *uuid, const char *prefix,
                        const ffb_read_result *r) {
    fprintf(fp,
            "%c%s%s  { \"filename\" : \"%s\", \"uuid\" : \"%s\", \"normalised_score\" : %f,  \"nblock\" : %zu,  "
            "\"sequence_length\" : %zu,  \"blocks_per_base\" : %f, \"nsample\" : %zu, \"trim\" : [ %zu, %zu ] }\n",
            lead, prefix, name, readname, uuid, -r->score / r->nblock, r->nblock, r->basecall_length,
            (float)r->nblock / (float)r->basecall_length, r->n, r->start, r->end);
}

void ffb_fprintf_read(enum ffb_outformat fmt, FILE *fp, const char *uuid, const char *readname, bool uuid_primary,
                      const char *prefix, const ffb_read_result *r) {
    if (!fp || !r || !r->basecall) return;
    const char *name = uuid_primary ? uuid : readname;
    switch (fmt) {
    case FFB_OUT_FASTA:
        header_line(fp, '>', name, readname, uuid, prefix, r);
        fputs(r->basecall, fp); fputc('\n', fp);
        break;
    case FFB_OUT_FASTQ:
        if (!r->quality) {
            fprintf(stderr, "Can't output fastq for reads without quality values\n");
            return;
        }
        header_line(fp, '@', name, readname, uuid, prefix, r);
        fputs(r->basecall, fp); fputs("\n+\n", fp);
        fputs(r->quality, fp); fputc('\n', fp);
        break;
    case FFB_OUT_SAM:
        /* unaligned record, then the sequence and quality once more on a line of their own: the reference prints
         * both (src/flappie_output.c:126-131) and a drop-in keeps its bytes */
        fprintf(fp, "%s%s\t4\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\n", prefix, name, r->basecall, r->quality ? r->quality : "");
        fputs(r->basecall, fp); fputc('\t', fp);
        if (r->quality) fputs(r->quality, fp);
        fputc('\n', fp);
        break;
    default:
        fprintf(stderr, "Invalid flappie output format\n");
        return;
    }
    fflush(fp);
}

void ffb_write_trace(FILE *fp, const char *name, const uint8_t *trace, size_t nblock, size_t nstate) {
    if (!fp || !name || !trace) return;
    const uint32_t len = (uint32_t)strlen(name), ns = (uint32_t)nstate;
    const uint64_t nrow = (uint64_t)nblock + 1;
    fwrite("FFBT", 1, 4, fp);
    fwrite(&len, sizeof len, 1, fp);
    fwrite(name, 1, len, fp);
    fwrite(&nrow, sizeof nrow, 1, fp);
    fwrite(&ns, sizeof ns, 1, fp);
    fwrite(trace, 1, (size_t)nrow * nstate, fp);
}
