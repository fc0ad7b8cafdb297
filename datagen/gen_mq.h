/* gen_mq.h -- MQ encoder state for datagen (reference encoder side; mqc.go:169-349). */
#ifndef GEN_MQ_H
#define GEN_MQ_H
#include <stdint.h>
enum { GEN_CTX_ZC0 = 0, GEN_CTX_SC0 = 9, GEN_CTX_MAG0 = 14, GEN_CTX_RL = 17, GEN_CTX_UNI = 18, GEN_NUM_CTX = 19 };
extern uint32_t gen_mq_qe[94];
extern uint8_t  gen_mq_nmps[94];
extern uint8_t  gen_mq_nlps[94];
void gen_mq_tables_init(void);
typedef struct {
    uint32_t A, C, CT;
    uint8_t *buf; int cap; int bp; int overflow;
    uint8_t ctx[GEN_NUM_CTX];
} gen_mqenc;
void gen_mqenc_init(gen_mqenc *e, uint8_t *buf, int cap);
void gen_mqenc_encode(gen_mqenc *e, int ctx, int d);
int  gen_mqenc_flush(gen_mqenc *e, const uint8_t **start);
#endif
